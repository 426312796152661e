// Variant probe: canonical recipe with knobs (rank, dtype, l2 promotion, dynamic smem, fence flavour).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int RANK, bool DYN, bool PROXYFENCE>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int bytes, float *out)
{
    __shared__ alignas(1024) float sbuf[DYN ? 1 : 2048];
    extern __shared__ __align__(1024) unsigned char dyn[];
    __shared__ alignas(8) uint64_t bar;
    float *dst = DYN ? reinterpret_cast<float *>(dyn) : sbuf;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
        if (PROXYFENCE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
        if (RANK == 2)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(s32(&bar)), "r"(c0), "r"(c1) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(s32(dst)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(s32(&bar)) : "memory");
    out[threadIdx.x] = dst[threadIdx.x];
}

typedef CUresult (*Fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                       const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main()
{
    const int W = 1024, H = 16, C = 8, N = 4;
    size_t n = (size_t)W * H * C * N;
    std::vector<float> h(n);
    for (size_t i = 0; i < n; i++) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 512);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void *sym; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
    Fn enc = (Fn)sym;
    auto run = [&](const char *name, int rank, bool dyn, bool pf, CUtensorMapL2promotion l2, CUtensorMapSwizzle sw) {
        CUtensorMap m{};
        cuuint64_t dims[4] = {W, H, C, N};
        cuuint64_t str[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
        cuuint32_t box2[2] = {32, 8}, box4[4] = {32, 1, 8, 1}, es[4] = {1, 1, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, dims, str, rank == 2 ? box2 : box4, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaError_t e;
        if (rank == 2 && !dyn && pf) k<2, false, true><<<1, 128>>>(m, 64, 3, 0, 0, 1024, o);
        else if (rank == 2 && !dyn && !pf) k<2, false, false><<<1, 128>>>(m, 64, 3, 0, 0, 1024, o);
        else if (rank == 2 && dyn) { cudaFuncSetAttribute(k<2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); k<2, true, false><<<1, 128, 65536>>>(m, 64, 3, 0, 0, 1024, o); }
        else if (rank == 4 && !dyn) k<4, false, false><<<1, 128>>>(m, 64, 3, 2, 1, 1024, o);
        else { cudaFuncSetAttribute(k<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); k<4, true, false><<<1, 128, 65536>>>(m, 64, 3, 2, 1, 1024, o); }
        e = cudaDeviceSynchronize();
        float res[2] = {-1, -1};
        if (e == cudaSuccess) cudaMemcpy(res, o, 8, cudaMemcpyDeviceToHost);
        printf("%-44s encode=%d kernel=%s first=%.0f\n", name, (int)r, cudaGetErrorString(e), res[0]);
        if (e != cudaSuccess) exit(0);
    };
    run("2d static proxyfence l2none", 2, false, true, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_SWIZZLE_NONE);
    run("2d static mbarfence  l2none", 2, false, false, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_SWIZZLE_NONE);
    run("2d static mbarfence  l2 128", 2, false, false, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_SWIZZLE_NONE);
    run("2d static mbarfence  l2 128 sw128", 2, false, false, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_SWIZZLE_128B);
    run("2d dynamic", 2, true, false, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_SWIZZLE_128B);
    run("4d static", 4, false, false, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_SWIZZLE_128B);
    run("4d dynamic", 4, true, false, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_SWIZZLE_128B);
    {   // NHWC-style: dims (C=32, W=64, H=4, N=4) box (32 ch, 128 w with element stride 2 -> 64 pixels, 1, 1), coordinate w odd
        CUtensorMap m{};
        cuuint64_t dims[4] = {32, 256, 4, 4};
        cuuint64_t str[3] = {32 * 4, 32 * 256 * 4, 32 * 256 * 4 * 4};
        cuuint32_t box[4] = {32, 128, 1, 1}, es[4] = {1, 2, 1, 1};
        CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cudaFuncSetAttribute(k<4, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        k<4, true, false><<<1, 128, 65536>>>(m, 0, 5, 1, 0, 32 * 64 * 4, o);
        cudaError_t e = cudaDeviceSynchronize();
        float res[64];
        if (e == cudaSuccess) cudaMemcpy(res, o, 256, cudaMemcpyDeviceToHost);
        printf("nhwc box w=128 estride 2, w0=5: encode=%d kernel=%s first=%.0f (expect %d) elem[32]=%.0f (expect swizzled row 1: pixel 7)\n", (int)r,
               cudaGetErrorString(e), res[0], (1 * 256 + 5) * 32, res[32]);
        if (e != cudaSuccess) return 0;
    }
    run_unaligned:
    {
        CUtensorMap m{};
        cuuint64_t dims[4] = {W, H, C, N};
        cuuint64_t str[3] = {W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * C * 4};
        cuuint32_t box4[4] = {32, 1, 8, 1}, es[4] = {1, 1, 1, 1};
        enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box4, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        k<4, true, false><<<1, 128, 65536>>>(m, 65, 3, 2, 1, 1024, o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("4d dynamic, inner coordinate 65 (unaligned): kernel=%s\n", cudaGetErrorString(e));
    }
    return 0;
}
