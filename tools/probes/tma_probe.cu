// tma_probe.cu -- standalone check of the 2D uint8 TMA box load used by k_score_tiled.
// usage: tma_probe <box_inner> <box_outer> <x0> <y0> <via_global_desc 0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void k(const __grid_constant__ CUtensorMap tmap, const CUtensorMap *gmap, int use_g, int c0, int c1, int bytes,
                  unsigned char *out)
{
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const CUtensorMap *m = use_g ? gmap : &tmap;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(smem_u32(sm)), "l"(m), "r"(c0), "r"(c1), "r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

typedef CUresult (*PFN)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                        const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    int bi = atoi(argv[1]), bo = atoi(argv[2]), x0 = atoi(argv[3]), y0 = atoi(argv[4]), useg = atoi(argv[5]);
    const int W = 1600, H = 1600;
    std::vector<unsigned char> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (unsigned char)((i * 2654435761u) >> 24);
    unsigned char *d, *dout; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    int bytes = bi * bo; cudaMalloc(&dout, bytes);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ce = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry point: %s fn=%p q=%d\n", cudaGetErrorString(ce), fn, (int)q);
    CUtensorMap map; memset(&map, 0, sizeof map);
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)W}; cuuint64_t strides[1] = {(cuuint64_t)H};
    cuuint32_t box[2] = {(cuuint32_t)bi, (cuuint32_t)bo}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((PFN)fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    CUtensorMap *gmap; cudaMalloc(&gmap, sizeof map); cudaMemcpy(gmap, &map, sizeof map, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    k<<<1, 256, bytes>>>(map, gmap, useg, y0, x0, bytes, dout);
    ce = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(ce));
    if (ce != cudaSuccess) return 1;
    std::vector<unsigned char> o(bytes); cudaMemcpy(o.data(), dout, bytes, cudaMemcpyDeviceToHost);
    long bad = 0;
    for (int r2 = 0; r2 < bo; r2++) for (int c = 0; c < bi; c++) {
        int gx = x0 + r2, gy = y0 + c;
        unsigned char want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[(size_t)gx * H + gy] : 0;
        if (o[(size_t)r2 * bi + c] != want) bad++;
    }
    printf("box %dx%d at (%d,%d) via %s: mismatches %ld\n", bi, bo, x0, y0, useg ? "global desc" : "grid_constant", bad);
    return bad != 0;
}
