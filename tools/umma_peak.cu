// umma_peak.cu - measures the issue-limited peak of tcgen05.mma on this GPU for the shapes the convolution kernels use:
// kind::tf32 and kind::f16, cta_group::1 (M=128) and cta_group::2 (M=256), N=192, operands resident in shared memory
// (no TMA, garbage data - only the rate matters).  One CTA (pair) per SM (pair), ITERS back-to-back MMAs with a commit
// every 4, as in the real kernels.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_peak tools/umma_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3fff);
    d |= (uint64_t)((16 >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int CTAS, bool TF32>
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t ad, uint64_t bd, uint32_t idesc)
{
    if constexpr (CTAS == 1 && TF32)
        asm volatile("{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    else if constexpr (CTAS == 1 && !TF32)
        asm volatile("{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    else if constexpr (CTAS == 2 && TF32)
        asm volatile("{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
    else
        asm volatile("{\n.reg .pred p;\nsetp.eq.b32 p, 0, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(ad), "l"(bd), "r"(idesc) : "memory");
}

template <int CTAS>
__device__ __forceinline__ void commit(uint64_t *bar)
{
    if constexpr (CTAS == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)1) : "memory");
}

template <int CTAS, bool TF32, int N>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int same_stage, int a_shift, int tma_bytes, const unsigned char *gsrc)
{
    extern __shared__ unsigned char raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint64_t tbar[4];
    __shared__ uint32_t slot;
    uint32_t rank = 0;
    if constexpr (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (8 * 16384 + 4 * 24576) / 4; i += blockDim.x) {     // plausible operand values
        uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
        float v = ((h >> 9) & 0xffff) / 65536.0f - 0.5f;
        if (TF32) reinterpret_cast<float *>(smem)[i] = v;
        else reinterpret_cast<uint32_t *>(smem)[i] = (__float_as_uint(v) >> 16) | (__float_as_uint(-v) & 0xffff0000u);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        for (int i = 0; i < 4; i++) mbar_init(&tbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (CTAS == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CTAS == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    // idesc: c f32 (1<<4); a/b format: tf32 = 2, bf16 = 1 at [7,10) and [10,13); N>>3 at 17; M>>4 at 24
    const uint32_t fmt = TF32 ? 2u : 1u;
    const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((CTAS == 2 ? 256 : 128) >> 4) << 24);
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 8 * 16384;
        const uint64_t hi = smem_desc(0) & 0xFFFFFFFF00000000ull;
        uint32_t phase = 0;
        for (int it = 0; it < iters; it += 4) {
            const int st = same_stage ? 0 : ((it >> 2) & 3);
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const uint64_t ad = hi | (uint64_t)(((a0 + st * 16384 + a_shift * 128 + kk * 32) >> 4) & 0x3fff);
                const uint64_t bd = hi | (uint64_t)(((b0 + st * 24576 + kk * 32) >> 4) & 0x3fff);
                mma<CTAS, TF32>(tmem, ad, bd, idesc);
            }
            if ((it & 63) == 60) {       // keep the queue bounded: wait for completion every 64 MMAs
                commit<CTAS>(&bar);
                mbar_wait(&bar, phase);
                phase ^= 1;
            }
        }
        commit<CTAS>(&bar);
        mbar_wait(&bar, phase);
    }
    if (warp == 0 && lane == 0 && tma_bytes > 0) {
        // background TMA traffic: tma_bytes per 4 MMAs into the four upper A stages (which the MMAs never read)
        uint32_t ph[4] = {0, 0, 0, 0};
        int slot = 0;
        const int n = iters / 4;
        for (int i = 0; i < n; i++) {
            if (i >= 4) { mbar_wait(&tbar[slot], ph[slot]); ph[slot] ^= 1; }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tbar[slot])), "r"(tma_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + (4 + slot) * 16384)),
                         "l"(gsrc + ((size_t)(blockIdx.x * 4 + slot) << 15)), "r"(tma_bytes), "r"(smem_u32(&tbar[slot])) : "memory");
            slot = (slot + 1) & 3;
        }
        for (int k = 0; k < 4 && k < n; k++) { mbar_wait(&tbar[slot], ph[slot]); ph[slot] ^= 1; slot = (slot + 1) & 3; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if constexpr (CTAS == 2) asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    else __syncthreads();
    if (warp == 1) {
        if constexpr (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256) : "memory");
    }
}

static void *g_src = nullptr;
template <int CTAS, bool TF32, int N>
void run(const char *name, int sms, int same_stage, int a_shift = 0, int tma_bytes = 0)
{
    const int iters = 1 << 16;
    const size_t smem = 1024 + 8 * 16384 + 4 * 24576;
    auto k = peak_kernel<CTAS, TF32, N>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(sms / CTAS * CTAS);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CTAS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, k, iters, same_stage, a_shift, tma_bytes, (const unsigned char *)g_src);
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const int K = TF32 ? 8 : 16;
    const double flops = 2.0 * 128 * N * K * (double)iters * (sms / CTAS * CTAS);
    printf("%-34s same_stage=%d shift=%d tma=%5d B/4MMA  %8.3f ms  %8.1f TFLOP/s  (%.1f clk/MMA at 1.9 GHz)\n", name, same_stage, a_shift, tma_bytes, best, flops / best / 1e9,
           best * 1e-3 * 1.9e9 / iters);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    cudaMalloc(&g_src, (size_t)148 * 4 << 15);
    cudaMemset(g_src, 0x3c, (size_t)148 * 4 << 15);
    // the upper four A stages receive the background TMA traffic, so MMAs rotate over stages 0..3 only
    run<1, true, 192>("tf32 cta_group::1 M128 N192 K8", sms, 0);
    run<2, true, 192>("tf32 cta_group::2 M256 N192 K8", sms, 0);
    for (int shift = 1; shift <= 9; shift += (shift == 3 ? 5 : 1)) run<2, true, 192>("tf32 cta_group::2 M256 N192 K8", sms, 0, shift);
    run<1, true, 192>("tf32 cta_group::1 M128 N192 K8", sms, 0, 1);
    for (int tb : {4096, 8192, 16384}) {
        run<1, true, 192>("tf32 cta_group::1 M128 N192 K8", sms, 0, 0, tb);
        run<2, true, 192>("tf32 cta_group::2 M256 N192 K8", sms, 0, 0, tb);
        run<2, true, 192>("tf32 cta_group::2 M256 N192 K8", sms, 0, 2, tb);
    }
    run<1, false, 192>("bf16 cta_group::1 M128 N192 K16", sms, 0);
    run<2, true, 96>("tf32 cta_group::2 M256 N96 K8", sms, 0);
    run<2, true, 256>("tf32 cta_group::2 M256 N256 K8", sms, 0);
    return 0;
}
