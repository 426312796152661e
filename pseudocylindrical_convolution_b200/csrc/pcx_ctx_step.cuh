// pcx_ctx_step.cuh - device building blocks shared by the two decoder step kernels (pcx_ctx.cu: one cooperative launch per
// wavefront step with grid barriers; pcx_flow.cu: one persistent dataflow kernel per decode): the GMM CDF row, the
// channels-last step network description, the layer-independent (cell, tap) geometry and the warp-per-cell masked
// convolution in the reference's reduction order.
#pragma once
#include "pcx_common.cuh"
#include <math.h>

namespace {

// ------------------------------------------------------------------------------------------------ GMM
// entropy_gmm_table_weight_kernel + _delta_kernel + _batch_forward_kernel + _check_kernel
// (extension/entropy_gmm_table_cuda.cu:29-56, :83-105, :136-153) in one launch, one thread per symbol.
// Expression shapes follow the reference's SASS: float v, FMUL s2*(v-mu), IEEE float division, erff,
// DFMA(erf, .5, .5), DFMA(f, w, ps) rounded to float per component, FMUL total*ps, DADD .5, truncation.
// w: mixture logits on entry, softmax weights on return; d: raw deltas on entry, clamped on return; c[0..nstep]: the table.
// NGC / NSC > 0 fix the mixture size / number of steps at compile time (every loop unrolls, the arrays live in registers);
// 0 takes the run-time value.  One source for both so that the expression shapes - and the bits - are the same.
// The row is built from three pieces so that a row can also be spread over the lanes of a thread group (gmm_boundary is the
// expensive part - a division, an erf and two double FMAs per mixture component): the persistent decoder kernel gives every
// boundary of a row its own lane.  gmm_cdf_row is the three pieces in sequence - one source, identical bits either way.
template <int NGC>
__device__ __forceinline__ void gmm_prepare(float *w, float *d, int ng_rt, float beta)
{
    const int ng = NGC > 0 ? NGC : ng_rt;
    float mval = -1e10f, psum = 0.f;
#pragma unroll
    for (int i = 0; i < ng; i++)
        if (mval < w[i]) mval = w[i];
#pragma unroll
    for (int i = 0; i < ng; i++) {
        w[i] = exp(w[i] - mval);
        psum += w[i];
    }
#pragma unroll
    for (int i = 0; i < ng; i++) {
        w[i] = w[i] / psum;
        float t = d[i];
        t = t < 0 ? beta : t + beta;
        d[i] = t;
    }
}
// c[pt], 1 <= pt < nstep, from the prepared weights / deltas
template <int NGC>
__device__ __forceinline__ float gmm_boundary(const float *w, const float *d, const float *mu, int ng_rt, int pt, float bias, float total,
                                              int form)
{
    const int ng = NGC > 0 ? NGC : ng_rt;
    const float s2 = (float)(1. / sqrt(2.0));
    float v = pt - 1 - bias + 0.5;
    float ps = 0;
    if (form == 0) {
        // entropy_gmm_table_batch_forward_kernel (:146-150): the whole term stays in double, one rounding per component
#pragma unroll
        for (int i = 0; i < ng; i++) {
            ps = ps + w[i] * (0.5 + 0.5 * erf(s2 * (v - mu[i]) / d[i]));
        }
    } else {
        // entropy_gmm_table_forward_kernel (:69-72): f is stored to float first, then a float FFMA
        float f;
#pragma unroll
        for (int i = 0; i < ng; i++) {
            f = 0.5 + 0.5 * erf(s2 * (v - mu[i]) / d[i]);
            ps = ps + w[i] * f;
        }
    }
    return (float)static_cast<int>(total * ps + 0.5);
}
// strict-monotonic fix-up (:83-105), on integer-valued floats as the reference does; c[0] and c[nstep] are set here
template <int NSC>
__device__ __forceinline__ void gmm_fixup(float *c, int nstep_rt, float total)
{
    const int nstep = NSC > 0 ? NSC : nstep_rt;
    c[0] = 0.f;
    c[nstep] = (float)static_cast<int>(total);
    float fb = 0.f, mv = 0.f;
    int midx = 0;
#pragma unroll
    for (int i = 0; i < nstep; i++) {
        if (c[i + 1] <= c[i]) fb += 1.f;
        c[i + 1] += fb;
        if (c[i + 1] - c[i] > mv) { mv = c[i + 1] - c[i]; midx = i; }
    }
    if (fb > 0.f) {
#pragma unroll
        for (int i = 0; i < nstep; i++)
            if (i >= midx) c[i + 1] -= fb;
    }
}
template <int NGC, int NSC>
__device__ __forceinline__ void gmm_cdf_row(float *w, float *d, const float *mu, int ng_rt, int nstep_rt, float bias, float total,
                                            float beta, int form, float *c)
{
    const int nstep = NSC > 0 ? NSC : nstep_rt;
    gmm_prepare<NGC>(w, d, ng_rt, beta);
#pragma unroll
    for (int pt = 1; pt < nstep; pt++) c[pt] = gmm_boundary<NGC>(w, d, mu, ng_rt, pt, bias, total, form);
    gmm_fixup<NSC>(c, nstep_rt, total);
}


struct StepLayer {
    const float *weight, *bias, *act, *add, *in;
    float *out;
    int gi, pad_out, constrain, cp_in, cp_out;
};
struct StepNet {
    int nlayers, nb, nimg, npart, G, h, W, pad, nstep, ng;
    float gmm_bias, gmm_total, gmm_beta, input_bias;
    Bands bands;
    const int *hband, *hrow, *hcol, *order;
    const float *htw;
    const int4 *cell;           // per entry of `order`: (col, global row, band, row in band) - no integer division on the hot path
    float halo_one;             // smallest float f with (double)f >= 1 - 1e-6: the "no causal source" test of pcx_ctx_pad_items
    float *sym_nchw;            // layers[0].in of the caller (NCHW, padded): receives symbol + input_bias for the Python side
    const float *prev;          // symbols decoded at the previous step, (image, cell of the window) - mapped host memory
    int *cdf;                   // CDF rows of this step, (image, cell of the window) x (nstep+1) - mapped host memory
    unsigned *bar;              // grid barrier counter
    unsigned long long *dbg;    // optional: globaltimer of block 0 at kernel entry, after every barrier and at exit (PCX_WAVE_TRACE)
    StepLayer L[PCX_WAVE_MAX_LAYERS];
};

__global__ void step_cellinfo_kernel(const int *__restrict__ order, int4 *__restrict__ cell, int n, int h, int W)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int hw = order[i], tw = hw % W, hp = hw / W;
    cell[i] = make_int4(tw, hp, hp / h, hp % h);
}

__device__ __forceinline__ void step_stamp(const StepNet &d, int step, int slot)
{
    if (d.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        d.dbg[(size_t)step * 32 + slot] = t;
    }
}

// Arrive with a release reduction, poll with RELAXED loads and fence once after the exit: an acquire load in the spin loop
// makes every poll invalidate the SM's L1 (CCTL.IVALL, 1.5 M per launch in the first profile) under the other resident block
// that is still computing on L1-cached activations.
__device__ __forceinline__ void grid_barrier(unsigned *bar, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        unsigned v;
        do {
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
        } while ((int)(v - target) < 0);
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
}

struct StepTap {
    const float *pa, *pb;       // mode 0: value = *pa; mode 1: lerp2(*pa, *pb, t); mode 2: lerp2(0, *pb, t); mode 3: 0
    float t;
    int mode;
};

// tap at row yi (relative to the band's first row, may be -pad..h+pad-1) and column xi (relative to the first valid column).
// The result is independent of the layer: every layer input is a channels-last plane set of the same padded geometry, so a
// source is identified by its CELL index (offa / offb, to be multiplied by the layer's channels per cell).
struct StepTapOff {
    int offa, offb;             // mode 0: value = in[offa]; mode 1: lerp2(in[offa], in[offb], t); mode 2: lerp2(0, in[offb], t); mode 3: 0;
                                // mode 4: lerp2(in[offa], 0, 1)
    float t;
    int mode;
};

__device__ __forceinline__ StepTapOff step_resolve_off(const StepNet &d, i64 pn, int g, int yi, int xi)
{
    StepTapOff r;
    r.offa = r.offb = 0;
    r.t = 0.f;
    r.mode = 3;
    const int h = d.h, W = d.W, pad = d.pad;
    const i64 ih = h + 2 * pad, iw = W + 2 * pad;
    const int wlg = d.bands.wl[g];
    if (xi < 0) return r;                                        // left pad stays 0 (entropy_context_cuda.cu:85-103)
    if (xi >= wlg) {                                             // right wrap: copy of the first pad columns of the same row
        const int ph = g * h + yi;
        if (xi >= wlg + pad || ph < 0 || ph >= h * d.npart) return r;
        xi -= wlg;
    }
    if (yi >= 0 && yi < h) {
        r.offa = r.offb = (int)(((pn * d.npart + g) * ih + yi + pad) * iw + xi + pad);
        r.mode = 0;
        return r;
    }
    const int s = yi < 0 ? 0 : 1, rr = yi < 0 ? yi + pad : yi - h;
    const int hr = (g * 2 + s) * pad + rr;
    const int pg = d.hband[hr];
    if (pg < 0) return r;                                        // pole rows stay 0
    const i64 e = (i64)hr * W + xi;
    const int q = d.hcol[e];
    const float t = d.htw[e];
    if (q < 0 && t >= d.halo_one) return r;                      // no causal source: the cell stays 0
    const i64 srow = ((pn * d.npart + pg) * ih + d.hrow[hr] + pad) * iw + pad;
    const int q1 = (q + 1 == d.bands.wl[pg]) ? 0 : q + 1;
    r.offa = (int)(srow + (q < 0 ? 0 : q));
    r.offb = (int)(srow + q1);
    r.t = t;
    // t == 1 ("left sample only", entropy_context_cuda.cu:146-158: q + 1 > tw): the right neighbour lies on a LATER plane and
    // does not exist yet when the reference fills this halo cell - it reads the zero-initialised buffer, fma(a, 1, 0 * 0).
    // Mode 4 reproduces exactly that without touching the neighbour (which the dataflow kernel would otherwise wait for).
    r.mode = q < 0 ? 2 : (t == 1.0f ? 4 : 1);
    return r;
}

__device__ __forceinline__ StepTap step_tap_of(const StepTapOff &o, const float *in, int cp)
{
    StepTap r;
    r.pa = in + (i64)o.offa * cp;
    r.pb = in + (i64)o.offb * cp;
    r.t = o.t;
    r.mode = o.mode;
    return r;
}

// NQ consecutive 128-bit loads from one base address in ONE asm statement (one predicate, immediate offsets, issued back to
// back; volatile keeps them ahead of the arithmetic that follows).  With on == 0 nothing is loaded and v is left untouched -
// the caller never consumes it.  L1-cached (.ca) on purpose: inside one launch every scratch buffer is written in exactly one
// phase and read only in later ones, behind a grid barrier with acquire semantics, and L1 starts clean at every launch - a
// cached line can never be stale, and the 5x5 windows of neighbouring cells (warps of the same block) overlap by two thirds.
__device__ __forceinline__ void ld_ca_v4x2(float4 (&v)[2], const float *p, int on)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.s32 p, %9, 0;\n"
        "@p ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%8];\n"
        "@p ld.global.ca.v4.f32 {%4, %5, %6, %7}, [%8+16];\n}"
        : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[0].z), "+f"(v[0].w), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[1].z), "+f"(v[1].w)
        : "l"(p), "r"(on)
        : "memory");
}
__device__ __forceinline__ void ld_ca_v4x6(float4 (&v)[6], const float *p, int on)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.s32 p, %25, 0;\n"
        "@p ld.global.ca.v4.f32 {%0, %1, %2, %3}, [%24];\n"
        "@p ld.global.ca.v4.f32 {%4, %5, %6, %7}, [%24+16];\n"
        "@p ld.global.ca.v4.f32 {%8, %9, %10, %11}, [%24+32];\n"
        "@p ld.global.ca.v4.f32 {%12, %13, %14, %15}, [%24+48];\n"
        "@p ld.global.ca.v4.f32 {%16, %17, %18, %19}, [%24+64];\n"
        "@p ld.global.ca.v4.f32 {%20, %21, %22, %23}, [%24+80];\n}"
        : "+f"(v[0].x), "+f"(v[0].y), "+f"(v[0].z), "+f"(v[0].w), "+f"(v[1].x), "+f"(v[1].y), "+f"(v[1].z), "+f"(v[1].w),
          "+f"(v[2].x), "+f"(v[2].y), "+f"(v[2].z), "+f"(v[2].w), "+f"(v[3].x), "+f"(v[3].y), "+f"(v[3].z), "+f"(v[3].w),
          "+f"(v[4].x), "+f"(v[4].y), "+f"(v[4].z), "+f"(v[4].w), "+f"(v[5].x), "+f"(v[5].y), "+f"(v[5].z), "+f"(v[5].w)
        : "l"(p), "r"(on)
        : "memory");
}
template <int NQ>
__device__ __forceinline__ void ld_ca_batch(float4 (&v)[NQ], const float *p, int on);
template <>
__device__ __forceinline__ void ld_ca_batch<2>(float4 (&v)[2], const float *p, int on) { ld_ca_v4x2(v, p, on); }
template <>
__device__ __forceinline__ void ld_ca_batch<6>(float4 (&v)[6], const float *p, int on) { ld_ca_v4x6(v, p, on); }
__device__ __forceinline__ void reg_fence_v4(float4 &a, float4 &b)
{
    asm volatile("" : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w), "+f"(b.x), "+f"(b.y), "+f"(b.z), "+f"(b.w));
}

// The block's share of a step: one (net image pn, plane) pair and a run of that plane's cells.  All cells of a plane carry
// the same channel group tc = step - plane, so the three output rows of (net, tc) are the only weights the block needs:
// they are staged in shared memory with cp.async one layer AHEAD (weights do not depend on the grid barrier).
struct StepChunk { int net, plane, cell0, ncell, img0, rem0; };   // cells cell0 .. cell0+ncell of the plane's (image, cell) list;
                                                                  // cell0 = img0 * cells_of_plane + rem0

#ifndef PCX_STEP_THREADS
#define PCX_STEP_THREADS 256
#endif
constexpr int STEP_THREADS = PCX_STEP_THREADS;
constexpr int STEP_MAX_RUNS = 32;  // runs of a block whose descriptors are cached in shared memory

__device__ __forceinline__ void cp_async4(float *smem_dst, const float *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// rows of outputs tc*3 .. tc*3+2 of net b, channel groups < gmax, interleaved as smem[(ci * 25 + tap) * 4 + og] so that one
// 128-bit shared load fetches the three weights of a (channel, tap); bias and PReLU slope sit behind them at 4 * wstride.
__device__ __forceinline__ void step_stage_weights(const StepNet &d, const StepLayer &l, int b, int tc, float *smem, int wstride)
{
    const int Ci = d.G * l.gi, Co = d.G * 3;
    int gmax = tc + 4 + (l.constrain == 6 ? 1 : 0);
    gmax = gmax > d.G ? d.G : gmax;
    const int nw = gmax * l.gi * 25;
    for (int i = threadIdx.x; i < nw * 3; i += blockDim.x) {
        const int og = i / nw, r = i % nw;
        cp_async4(smem + r * 4 + og, l.weight + (((i64)b * Co + tc * 3 + og) * Ci) * 25 + r);
    }
    if (threadIdx.x < 3) cp_async4(smem + 4 * wstride + threadIdx.x, l.bias + b * Co + tc * 3 + threadIdx.x);
    else if (threadIdx.x < 6 && l.act != nullptr) cp_async4(smem + 4 * wstride + threadIdx.x, l.act + b * Co + tc * 3 + threadIdx.x - 3);
}

// ---- dataflow synchronisation of the persistent decoder kernel (pcx_flow.cu)
constexpr unsigned FLOW_SENTINEL = 0xFFFFBEEFu;      // negative NaN with a payload; FADD / FMUL / FFMA only ever produce 0x7fffffff
struct FlowCtl {
    unsigned *abort_flag;                             // device memory: != 0 -> stop waiting (host error or time-out), results are void
    unsigned long long timeout_ns;                    // a single wait longer than this sets abort_flag = 2
};
__device__ __forceinline__ float flow_ld(const float *p)
{
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return __uint_as_float(v);
}
// spin on one scalar in L2 until its producer has replaced the sentinel
__device__ __noinline__ float flow_poll(const float *p, FlowCtl *ctl)
{
    unsigned v, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
        // back-to-back polling on purpose: a __nanosleep back-off between the polls was measured and costs far more in late
        // detection (20.0 -> 22.0 ms for one 512x1024 image, 75 -> 114 ms for eight) than it saves in L2 traffic
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
        if (v != FLOW_SENTINEL) break;
        if ((++spins & 127u) == 0) {
            unsigned a;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(a) : "l"(ctl->abort_flag) : "memory");
            if (a != 0) break;
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > ctl->timeout_ns) { atomicExch(ctl->abort_flag, 2u); break; }
        }
    }
    return __uint_as_float(v);
}
// One warp per cell.  Lane = filter tap (kh, kw) (25 live lanes); it runs the GI chains of its tap - the reference's virtual
// lanes m*25 + tap - over the allowed channel groups in ascending order, 8 groups per batch of 128-bit loads.  The chain sums
// are then moved to the virtual-lane positions (lane t takes lanes t, t+32, t+64 of the reference block) and folded exactly
// like the reference: [t]+=[t+64], [t]+=[t+32], shuffle-down 16..1.
// Per-step tap cache (shared memory).  The geometry of a (cell, tap) - which band / row / column it reads, through which halo
// table entries - does not depend on the layer, but resolving it costs three dependent L2 round trips (plane prefix -> cell
// record -> halo tables) that used to be paid again in every one of the 12 layers, right on the latency-bound critical path of
// a step.  The first STEP_CACHE_CELLS cells of a block's work list are resolved ONCE, before the first grid barrier.
constexpr int STEP_CACHE_CELLS = 64;
struct StepCellRec { int pn, g, th, tw; };

//
// FLOW (pcx_flow.cu): no grid barrier separates the layers.  Every scratch scalar is written exactly once per decode into a
// buffer pre-filled with FLOW_SENTINEL (a NaN payload no arithmetic instruction produces), so a value is its own "ready" flag.
// Of a chain's inputs only the LAST allowed channel group can belong to the current step (k + plane = step, see the mask
// rule above; earlier groups were final when the step began): that element - and the residual source - is validated and, if
// still the sentinel, polled from L2 until its producer (another warp, usually another block) has stored it.
template <int GI, bool FLOW = false>
__device__ __forceinline__ void step_conv_phase(const StepNet &d, const StepLayer &l, int step, const StepChunk &ch, const int *start,
                                                const float *ws, int wstride, const StepTapOff *tap_cache, const StepCellRec *cell_cache,
                                                int cache_base, int img_base = 0, FlowCtl *ctl = nullptr, int cache_cells = STEP_CACHE_CELLS)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int G = d.G, h = d.h, W = d.W;
    const int cp = l.cp_in;
    constexpr int NQ = 2 * GI;                                  // float4 per batch of 8 channel groups
    const int c6 = l.constrain == 6 ? 1 : 0;
    const int tc = step - ch.plane;
    const int first = start[ch.plane], pcells = start[ch.plane + 1] - start[ch.plane];
    const bool live = lane < 25;
    const int kw = lane % 5, kh = (lane / 5) % 5;
    int nk_tap = live ? tc + 4 - kh - kw + c6 : 0;
    nk_tap = nk_tap > G ? G : (nk_tap < 0 ? 0 : nk_tap);
    int gmax = tc + 4 + c6;                                     // longest chain of the block (tap 0,0)
    gmax = gmax > G ? G : gmax;
    const float4 *wl_ = reinterpret_cast<const float4 *>(ws) + kh * 5 + kw;
    for (int k = warp; k < ch.ncell; k += nwarp) {
#ifdef PCX_FLOW_PROBE
        const bool probe = FLOW && d.dbg != nullptr && blockIdx.x == d.nimg && warp == 0;      // block 1 of image 0, warp 0
        long long pt[6] = {0, 0, 0, 0, 0, 0};
        if (probe) { __syncwarp(); pt[0] = clock64(); }
#define PCX_PROBE(i) if (probe) { __syncwarp(); pt[i] = clock64(); }
#else
#define PCX_PROBE(i)
#endif
        int pn, tw, g, th;
        StepTap tp;
        tp.pa = tp.pb = l.in;
        tp.t = 0.f;
        tp.mode = 3;
        if (cache_base + k < cache_cells) {
            const StepCellRec cr = cell_cache[cache_base + k];
            pn = cr.pn; g = cr.g; th = cr.th; tw = cr.tw;
            if (live) tp = step_tap_of(tap_cache[(cache_base + k) * 25 + lane], l.in, cp);
        } else {
            int img = ch.img0, ce = ch.rem0 + k;                 // (image, cell of the plane) of entry cell0 + k
            while (ce >= pcells) { ce -= pcells; img++; }
            pn = ch.net * d.nimg + img_base + img;
            const int4 ci = d.cell[first + ce];
            tw = ci.x; g = ci.z; th = ci.w;
            if (live) tp = step_tap_of(step_resolve_off(d, pn, g, th + kh - 2, tw + kw - 2), l.in, cp);
        }
        const int nk = tp.mode == 3 ? 0 : nk_tap;               // a zero input leaves the chains at +0.0f
        // residual source of the three outputs: final since two phases ago, fetched now so that the epilogue waits for nothing
        const i64 o = ((((i64)pn * d.npart + g) * (h + 2 * l.pad_out) + th + l.pad_out) * (W + 2 * l.pad_out) + tw + l.pad_out) * l.cp_out + tc * 3;
        float addv[3] = {0.f, 0.f, 0.f};
        if (l.add != nullptr && lane < 3) addv[0] = FLOW ? flow_ld(l.add + o + lane) : __ldcg(l.add + o + lane);
        float acc[GI][3];
#pragma unroll
        for (int m = 0; m < GI; m++) acc[m][0] = acc[m][1] = acc[m][2] = 0.f;
        if (FLOW) {
            // The chain's LAST element (the only one that can belong to the current step) is fetched on its own, straight from
            // L2, together with the batches of the earlier elements - all loads of the task's first round trip in flight at once.
            // (The first version validated the loaded batches in place: ~150 compare / select instructions per batch to find one
            // run-time-indexed element in registers - 45 % of a task's cycles at the ~10 cycles per instruction of this
            // latency-bound kernel, measured with clock64.)  The prefix k < nk - 1 runs through the batched loop, the last FFMA
            // of every chain follows it: same ascending order.
            float la[GI], lb[GI];
#pragma unroll
            for (int m = 0; m < GI; m++) la[m] = lb[m] = 0.f;
            const bool use_a = nk > 0 && tp.mode != 2, use_b = nk > 0 && (tp.mode == 1 || tp.mode == 2);
            if (use_a) {
#pragma unroll
                for (int m = 0; m < GI; m++) la[m] = flow_ld(tp.pa + (nk - 1) * GI + m);
            }
            if (use_b) {
#pragma unroll
                for (int m = 0; m < GI; m++) lb[m] = flow_ld(tp.pb + (nk - 1) * GI + m);
            }
            const int np = nk - 1;                                // elements of the batched prefix
            PCX_PROBE(1)
            // (A two-stage pipeline - the loads of batch i + 1 issued before batch i's FFMAs - cuts the prefix phase from 5000 to
            // 4000 cycles per task in the clock64 probe below, but needs 199 registers: one block per SM, or 200 bytes of spills at
            // 128 registers - 64.5 -> 106 / 85.7 ms for 8 images.  Dropped.)
            for (int c0 = 0; c0 < gmax - 1; c0 += 8) {
                const int on = np > c0;
                float4 xa[NQ], xb[NQ];
#pragma unroll
                for (int q = 0; q < NQ; q++) xa[q] = xb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                ld_ca_batch<NQ>(xa, tp.pa + c0 * GI, on && tp.mode != 2);
                ld_ca_batch<NQ>(xb, tp.pb + c0 * GI, on && (tp.mode == 1 || tp.mode == 2));
                float va[8 * GI];
#pragma unroll
                for (int q = 0; q < NQ; q++) { va[4 * q] = xa[q].x; va[4 * q + 1] = xa[q].y; va[4 * q + 2] = xa[q].z; va[4 * q + 3] = xa[q].w; }
                if (tp.mode != 0) {                              // halo / wrap-of-halo tap: causal 2-tap interpolation (+0 when mode 2 / 4)
                    float vb[8 * GI];
#pragma unroll
                    for (int q = 0; q < NQ; q++) { vb[4 * q] = xb[q].x; vb[4 * q + 1] = xb[q].y; vb[4 * q + 2] = xb[q].z; vb[4 * q + 3] = xb[q].w; }
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        if (c0 + u < np) {
#pragma unroll
                            for (int m = 0; m < GI; m++) va[u * GI + m] = lerp2_ref(va[u * GI + m], vb[u * GI + m], tp.t);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    if (c0 + u < np) {
#pragma unroll
                        for (int m = 0; m < GI; m++) {
                            const float4 w = wl_[((c0 + u) * GI + m) * 25];
                            const float v = va[u * GI + m];
                            acc[m][0] = __fmaf_rn(v, w.x, acc[m][0]);
                            acc[m][1] = __fmaf_rn(v, w.y, acc[m][1]);
                            acc[m][2] = __fmaf_rn(v, w.z, acc[m][2]);
                        }
                    }
                }
            }
            PCX_PROBE(2)
            if (nk > 0) {
#pragma unroll
                for (int m = 0; m < GI; m++) {
                    if (use_a && __float_as_uint(la[m]) == FLOW_SENTINEL) la[m] = flow_poll(tp.pa + np * GI + m, ctl);
                    if (use_b && __float_as_uint(lb[m]) == FLOW_SENTINEL) lb[m] = flow_poll(tp.pb + np * GI + m, ctl);
                }
            }
            PCX_PROBE(3)
            if (nk > 0) {
#pragma unroll
                for (int m = 0; m < GI; m++) {
                    const float v = tp.mode == 0 ? la[m] : lerp2_ref(la[m], lb[m], tp.t);
                    const float4 w = wl_[(np * GI + m) * 25];
                    acc[m][0] = __fmaf_rn(v, w.x, acc[m][0]);
                    acc[m][1] = __fmaf_rn(v, w.y, acc[m][1]);
                    acc[m][2] = __fmaf_rn(v, w.z, acc[m][2]);
                }
            }
        } else
        for (int c0 = 0; c0 < gmax; c0 += 8) {
            const int on = nk > c0;
            float4 xa[NQ], xb[NQ];
#pragma unroll
            for (int q = 0; q < NQ; q++) xa[q] = xb[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            // both batches of 128-bit loads are in flight before the first FFMA can wait on one
            ld_ca_batch<NQ>(xa, tp.pa + c0 * GI, on && tp.mode != 2);
            ld_ca_batch<NQ>(xb, tp.pb + c0 * GI, on && (tp.mode == 1 || tp.mode == 2));
            float va[8 * GI];
#pragma unroll
            for (int q = 0; q < NQ; q++) { va[4 * q] = xa[q].x; va[4 * q + 1] = xa[q].y; va[4 * q + 2] = xa[q].z; va[4 * q + 3] = xa[q].w; }
            if (tp.mode != 0) {                                  // halo / wrap-of-halo tap: causal 2-tap interpolation (+0 when mode 2)
                float vb[8 * GI];
#pragma unroll
                for (int q = 0; q < NQ; q++) { vb[4 * q] = xb[q].x; vb[4 * q + 1] = xb[q].y; vb[4 * q + 2] = xb[q].z; vb[4 * q + 3] = xb[q].w; }
#pragma unroll
                for (int e = 0; e < 8 * GI; e++) va[e] = lerp2_ref(va[e], vb[e], tp.t);
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                if (c0 + u < nk) {
#pragma unroll
                    for (int m = 0; m < GI; m++) {
                        const float4 w = wl_[((c0 + u) * GI + m) * 25];
                        const float v = va[u * GI + m];
                        acc[m][0] = __fmaf_rn(v, w.x, acc[m][0]);
                        acc[m][1] = __fmaf_rn(v, w.y, acc[m][1]);
                        acc[m][2] = __fmaf_rn(v, w.z, acc[m][2]);
                    }
                }
            }
        }
        PCX_PROBE(4)
        // chain (m, tap) = reference lane i = m*25 + tap; lane t now collects lanes t, t+32, t+64
        float sum[3];
#pragma unroll
        for (int og = 0; og < 3; og++) {
            float v0, v1 = 0.f, v2 = 0.f;
            if (GI == 1) {
                v0 = acc[0][og];                                // lanes >= 25 hold +0
            } else {
                const int i0 = lane, i1 = lane + 32, i2 = lane + 64;
                const float a0 = __shfl_sync(0xffffffffu, acc[0][og], i0 % 25);
                const float a1 = __shfl_sync(0xffffffffu, acc[GI > 1 ? 1 : 0][og], i0 % 25);
                v0 = i0 < 25 ? a0 : a1;
                const float b1 = __shfl_sync(0xffffffffu, acc[GI > 1 ? 1 : 0][og], i1 % 25);
                const float b2 = __shfl_sync(0xffffffffu, acc[GI > 2 ? 2 : 0][og], i1 % 25);
                v1 = i1 < 50 ? b1 : b2;
                const float c2 = __shfl_sync(0xffffffffu, acc[GI > 2 ? 2 : 0][og], i2 % 25);
                v2 = i2 < 75 ? c2 : 0.f;
            }
            const float s0 = __fadd_rn(v0, v2);                 // [t] += [t+64]
            const float s1 = __fadd_rn(v1, 0.f);
            sum[og] = __fadd_rn(s0, s1);                        // [t] += [t+32]
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
#pragma unroll
            for (int og = 0; og < 3; og++) sum[og] = __fadd_rn(sum[og], __shfl_down_sync(0xffffffffu, sum[og], off));
        if (FLOW && l.add != nullptr && lane < 3 && __float_as_uint(addv[0]) == FLOW_SENTINEL) addv[0] = flow_poll(l.add + o + lane, ctl);
        addv[1] = __shfl_sync(0xffffffffu, addv[0], 1);
        addv[2] = __shfl_sync(0xffffffffu, addv[0], 2);
        if (lane == 0) {
#pragma unroll
            for (int og = 0; og < 3; og++) {
                float v = __fadd_rn(sum[og], ws[4 * wstride + og]);
                if (l.act != nullptr && v < 0.f) v = __fmul_rn(v, ws[4 * wstride + 3 + og]);
                if (l.add != nullptr) v = __fadd_rn(v, addv[og]);
                if (FLOW) __stcg(l.out + o + og, v);          // to L2, where the consumers poll
                else l.out[o + og] = v;
            }
        }
#ifdef PCX_FLOW_PROBE
        if (probe) {
            __syncwarp();
            const long long t = clock64();
            if (lane == 0) {       // cycles: setup + issue | prefix (load latency + FMAs) | wait for the last elements | last FMAs | fold + store
                atomicAdd(d.dbg + 0, (unsigned long long)(pt[1] - pt[0]));
                atomicAdd(d.dbg + 1, (unsigned long long)(pt[2] - pt[1]));
                atomicAdd(d.dbg + 2, (unsigned long long)(pt[3] ? pt[3] - pt[2] : 0));
                atomicAdd(d.dbg + 3, (unsigned long long)(pt[4] - (pt[3] ? pt[3] : pt[2])));
                atomicAdd(d.dbg + 4, (unsigned long long)(t - pt[4]));
                atomicAdd(d.dbg + 5, 1ull);
            }
        }
#endif
    }
}

// run `id` of the step: for every (net, plane) pair of the window the cells of ALL images (they share the weights) form one
// list of nimg * cells entries, cut into runs of at most S
__device__ __forceinline__ StepChunk step_chunk_of(const int *__restrict__ start, int id, int p0, int np, int S, int nb, int nimg)
{
    StepChunk c = {0, p0, 0, 0, 0, 0};
    for (int q = p0; q < p0 + np; q++) {
        const int pc = start[q + 1] - start[q];
        const int cells = pc * nimg;
        const int nrun = (cells + S - 1) / S;
        if (id < nrun * nb) {
            const int run = id % nrun;
            c.net = id / nrun; c.plane = q; c.cell0 = run * S;
            c.ncell = cells - run * S < S ? cells - run * S : S;
            c.img0 = c.cell0 / pc;
            c.rem0 = c.cell0 - c.img0 * pc;
            return c;
        }
        id -= nrun * nb;
    }
    return c;
}


}  // namespace
