// tma.cu -- host side of tma.cuh: tensor-map encoding through the driver entry point.
#include <cudaTypedefs.h>

#include "tma.cuh"

namespace fen {

static PFN_cuTensorMapEncodeTiled encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    }
    return fn;
}

int tma_encode_field(const Layout& L, const double* base, int box_x, int box_y, CUtensorMap* out) {
    PFN_cuTensorMapEncodeTiled fn = encode_fn();
    if (!fn) return set_error(FEN_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)L.px, (cuuint64_t)L.ny + 2, (cuuint64_t)L.nzl + 2};
    const cuuint64_t strides[2] = {(cuuint64_t)L.sy * sizeof(double), (cuuint64_t)L.sz * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return set_error(FEN_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (box %d x %d, pitch %d)", (int)r,
                         box_x, box_y, L.px);
    return FEN_OK;
}

int field_tmap(fen_ctx* c, const double* base, int box_x, int box_y, CUtensorMap** out) {
    const TmapKey key{base, box_x, box_y};
    auto it = c->tmaps.find(key);
    if (it == c->tmaps.end()) {
        CUtensorMap m;
        FEN_TRY(tma_encode_field(c->L, base, box_x, box_y, &m));
        it = c->tmaps.emplace(key, m).first;
    }
    *out = &it->second;
    return FEN_OK;
}

}  // namespace fen
