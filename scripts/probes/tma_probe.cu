// Probe: which (x0, box) combinations does a 3-D fp64 TMA tile load accept on this part?
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, int x0, int y0, int z, int nbox, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(nbox * 8) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(x0), "r"(y0), "r"(z), "r"(smem_u32(bar))
            : "memory");
    }
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    const double* s = reinterpret_cast<const double*>(smem);
    for (int e = threadIdx.x; e < nbox; e += blockDim.x) out[e] = s[e];
}

int main() {
    const int px = 96, ny2 = 20, nz2 = 6;
    std::vector<double> h((size_t)px * ny2 * nz2);
    for (size_t e = 0; e < h.size(); ++e) h[e] = (double)e;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8);
    cudaMalloc(&o, 16384);
    cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 + 64);
    const int cases[][3] = {{64, 10, 16}, {66, 10, 16}, {66, 10, 14}, {68, 10, 14}, {68, 10, 78}, {68, 10, 2}};
    for (auto& cs : cases) {
        const int bx = cs[0], by = cs[1], x0 = cs[2];
        CUtensorMap m;
        const cuuint64_t dims[3] = {px, ny2, nz2};
        const cuuint64_t str[2] = {px * 8, (cuuint64_t)px * ny2 * 8};
        const cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1};
        const cuuint32_t es[3] = {1, 1, 1};
        CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box %dx%d x0 %d: encode failed %d\n", bx, by, x0, (int)r); continue; }
        probe<<<1, 128, 16384 + 64>>>(m, x0, 3, 2, bx * by, o);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("box %dx%d x0 %d: kernel error: %s\n", bx, by, x0, cudaGetErrorString(e)); return 1; }
        std::vector<double> g((size_t)bx * by);
        cudaMemcpy(g.data(), o, g.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int yy = 0; yy < by; ++yy)
            for (int xx = 0; xx < bx; ++xx) {
                const int gx = x0 + xx, gy = 3 + yy;
                double want = (gx < px && gy < ny2) ? h[(size_t)gx + (size_t)px * (gy + (size_t)ny2 * 2)] : 0.0;
                if (g[(size_t)yy * bx + xx] != want) ++bad;
            }
        printf("box %dx%d x0 %d: ok, %d mismatches\n", bx, by, x0, bad);
    }
    return 0;
}
