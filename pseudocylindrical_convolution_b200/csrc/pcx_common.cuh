// pcx_common.cuh - shared host/device helpers for libpcx (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include "../../include/pcx.h"

typedef long long i64;

// ---------------------------------------------------------------------------------------------- errors
void pcx_set_error(const char *fmt, ...);
void pcx_append_error(const char *fmt, ...);     // adds context to the message of the error being returned
extern std::atomic<long long> g_pcx_launches;

#define PCX_REQUIRE(cond, ...)                                   \
    do {                                                         \
        if (!(cond)) {                                           \
            pcx_set_error(__VA_ARGS__);                          \
            return PCX_EINVAL;                                   \
        }                                                        \
    } while (0)

#define PCX_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            pcx_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return PCX_ECUDA;                                                                 \
        }                                                                                     \
    } while (0)

// every kernel launch goes through this: counts launches (bench.py gpu_launches) and checks the launch
#define PCX_LAUNCHED()                                                                   \
    do {                                                                                 \
        g_pcx_launches.fetch_add(1, std::memory_order_relaxed);                          \
        cudaError_t e__ = cudaGetLastError();                                            \
        if (e__ != cudaSuccess) {                                                        \
            pcx_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return PCX_ECUDA;                                                            \
        }                                                                                \
    } while (0)

// ---------------------------------------------------------------------------------------------- bands
struct Bands {
    int npart;
    int wl[PCX_MAX_PART];
};

static inline int make_bands(Bands &b, const int *wl, int npart)
{
    if (npart < 1 || npart > PCX_MAX_PART || wl == nullptr) return -1;
    b.npart = npart;
    for (int i = 0; i < PCX_MAX_PART; i++) b.wl[i] = i < npart ? wl[i] : 0;
    return 0;
}

static inline int ceil_div(i64 a, i64 b) { return (int)((a + b - 1) / b); }

int pcx_sm_count();

// "first call on this device": function attributes (dynamic shared memory limits, occupancy figures) are per device, and the
// reference API allows several devices in one process.  PCX_ONCE_PER_DEVICE(flag) { ... } runs its body once for every device
// it is reached on; `flag` is a function-local static PcxDeviceOnce.  The body runs under the flag's mutex, so a second host
// thread cannot launch before the attribute is set.
#include <mutex>
struct PcxDeviceOnce {
    std::mutex mu;
    unsigned long long done = 0;
    int slot[64] = {};                       // a per-device value computed by the body (e.g. an occupancy figure)
};
static inline int pcx_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    return dev;
}
#define PCX_ONCE_PER_DEVICE(flag)                                                                     \
    for (struct { std::unique_lock<std::mutex> lk; int dev; bool go; } o__ = {std::unique_lock<std::mutex>((flag).mu), pcx_current_device(), true}; \
         o__.go && !(((flag).done >> o__.dev) & 1ull); (flag).done |= 1ull << o__.dev, o__.go = false)

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------- PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA bulk copy global -> shared (16-byte aligned, size multiple of 16), completes on an mbarrier.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
// streaming 128-bit store (output tiles are written once and not re-read by the same kernel)
__device__ __forceinline__ void st_cs_f4(float *p, float4 v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_cs_f2(float *p, float2 v)
{
    asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// 2-tap halo interpolation in the reference's compiled shape (pseudo_pad.cu:77): r=1-t; r=b*r; fma(a,t,r)
__device__ __forceinline__ float lerp2_ref(float a, float b, float t)
{
    float r = __fsub_rn(1.0f, t);
    r = __fmul_rn(b, r);
    return __fmaf_rn(a, t, r);
}
// 4-tap cubic in the reference's compiled shapes.  nvcc picks which product of p1*i1 + p2*i2 + p3*i3 + p4*i4 is
// the plain FMUL (the rest are chained FFMAs), and it picks differently in the two kernels (SASS, nvcc 12.9):
//   sphere_slice_forward_kernel  (sphere_slice_cuda.cu:109-113): FMUL p2*i2, FFMA p1*i1, FFMA p3*i3, FFMA p4*i4
//   sphere_uslice_forward_kernel (sphere_uslice_cuda.cu:91-96) : FMUL p1*i1, FFMA p2*i2, FFMA p3*i3, FFMA p4*i4
template <bool SLICE_ORDER>
__device__ __forceinline__ float tap4_ref(float4 w, float a, float b, float c, float d)
{
    float r;
    if (SLICE_ORDER) {
        r = __fmul_rn(w.y, b);
        r = __fmaf_rn(w.x, a, r);
    } else {
        r = __fmul_rn(w.x, a);
        r = __fmaf_rn(w.y, b, r);
    }
    r = __fmaf_rn(w.z, c, r);
    r = __fmaf_rn(w.w, d, r);
    return r;
}
#endif
