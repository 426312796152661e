// Stand-alone probe: one TMA tensor load with the same tensor-map recipe as pcx_conv_tc.cu (debug helper).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap map, float *out, int c0, int c1, int c2, int c3, int nfloats)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(nfloats * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(smem_u32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main(int argc, char **argv)
{
    int pitch = 132, H = 10, C = 96, N = 16;
    int bw = argc > 1 ? atoi(argv[1]) : 32, bc = argc > 2 ? atoi(argv[2]) : 32, es = argc > 3 ? atoi(argv[3]) : 1;
    int swz = argc > 4 ? atoi(argv[4]) : 3;
    size_t n = (size_t)pitch * H * C * N;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)(i % 100003);
    float *d, *o;
    cudaMalloc(&d, n * 4);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    int nfl = (bw / es) * bc;
    cudaMalloc(&o, nfl * 4);
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                           const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Fn enc = (Fn)sym;
    CUtensorMap m;
    cuuint64_t dims[4] = {(cuuint64_t)pitch, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * H * 4, (cuuint64_t)pitch * H * C * 4};
    cuuint32_t box[4] = {(cuuint32_t)bw, 1, (cuuint32_t)bc, 1};
    cuuint32_t estr[4] = {(cuuint32_t)es, 1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     (CUtensorMapSwizzle)swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rc=%d (query %d) box w=%d c=%d estride=%d swizzle=%d\n", (int)r, (int)q, bw, bc, es, swz);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    probe<<<1, 128, 65536>>>(m, o, 3, 2, 32, 1, nfl);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> res(nfl);
        cudaMemcpy(res.data(), o, nfl * 4, cudaMemcpyDeviceToHost);
        // expected element (w=3+i*es, h=2, c=32+j, n=1)
        size_t base = ((size_t)1 * C + 32) * H * pitch + 2 * pitch + 3;
        printf("first row:");
        for (int i = 0; i < 8; i++) printf(" %.0f", res[i]);
        printf("\nexpect   :");
        for (int i = 0; i < 8; i++) printf(" %.0f", h[base + (size_t)i * es]);
        printf("\nrow1 (c+1) first: %.0f expect(unswizzled pos) %.0f\n", res[bw / es], h[base + (size_t)H * pitch]);
    }
    return 0;
}
