// tma.cuh -- thin wrappers over the sm_100a bulk-tensor copy engine (TMA) and mbarriers used by the
// stencil kernels to stage (tile + halo) boxes of a field in shared memory.
//
// Host: a CUtensorMap describes a field in its padded device layout (fen_internal.cuh: px x (ny+2) x
// (nzl+2) doubles) with a box of (BX, BY, 1) elements; it is encoded with cuTensorMapEncodeTiled, fetched
// through cudaGetDriverEntryPoint so that libfen_gpu.so does not link against libcuda.
// Device: one thread arms an mbarrier with the byte count and issues cp.async.bulk.tensor.3d; every thread
// that reads the tile waits on the barrier's phase parity.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fen_internal.cuh"

namespace fen {

// encodes the 3-D tiled map of a field; returns FEN_OK or sets the error
int tma_encode_field(const Layout& L, const double* base, int box_x, int box_y, CUtensorMap* out);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
// box (x, y, z) of the tensor `map` -> shared memory at dst, completion counted on bar
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
#endif

}  // namespace fen
