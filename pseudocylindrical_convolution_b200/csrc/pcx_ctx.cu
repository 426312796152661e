// pcx_ctx.cu - the masked autoregressive context model evaluated along the 3-D wavefront
// row + column + channel-group = step, and the GMM integer CDF tables that feed the arithmetic coder.
//
// Every output scalar is reduced in the reference's order: 128 virtual lanes, lane
// i < 25*gi owns tap (kw = i%5, kh = (i/5)%5, m = i/25) and chains FFMAs over the allowed channel groups in
// ascending order; lanes fold [t]+=[t+64], [t]+=[t+32], then shuffle-down 16..1.  Here one WARP produces one
// scalar: lane l carries virtual lanes l, l+32, l+64 in registers, so the folds are register adds and the
// tail is the same shuffle tree - bit-identical results with a quarter of the threads and no shared memory.
#include "pcx_common.cuh"
#include "pcx_flow.h"
#include <stdlib.h>
#include <atomic>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <string.h>
#include <math.h>
#include <vector>

namespace {

struct Window { int first, count; };

__device__ __forceinline__ void cp_async4_ctx(float *smem_dst, const float *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

// planes [max(0, s-G+1), min(s+1, Hf+W-1)) of the wavefront order (entropy_conv_cuda_v2.cu:389-391)
inline Window wave_window(const int *h_start, int psum, int G, int Hf, int W)
{
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W - 2 ? psum + 1 : Hf + W - 1;
    if (st > en) st = en;
    return {h_start[st], h_start[en] - h_start[st]};
}

inline int grid_for(i64 total, int threads)
{
    i64 want = (total + threads - 1) / threads;
    return (int)(want < 1 ? 1 : (want > 0x7fffffff ? 0x7fffffff : want));
}

// ------------------------------------------------------------------------------------------------ ctx pad
// entropy_ctx_pad_run2_forward_kernel (extension/entropy_ctx_pad_run2_cuda.cu:33-65)
__global__ void ctx_pad_kernel(float *__restrict__ buf, Bands bands, const int *__restrict__ hband, const int *__restrict__ hrow,
                               const int *__restrict__ hcol, const float *__restrict__ htw, const int4 *__restrict__ items,
                               int first, int nitems, int nrep, int cpn, int C, int h, int W, int pad, int psum, int full, int kind)
{
    // full == 0: the items of the current plane window, channel group psum - plane (one wavefront step).
    // full == 1: every channel group of every item (one-shot encoder); kind selects halo (0) or right-wrap (1) items,
    //            because a wrap item of a halo row copies a halo cell and must run after it.
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 total = (i64)nitems * cpn * nrep * (full ? C / cpn : 1);
    if (idx >= total) return;
    int it, cc, grp_full = 0;
    i64 n;
    const int G = C / cpn;
    if (total < (1ll << 31)) {                         // the usual case: 32-bit divisions (a 64-bit one costs ~100 instructions)
        const unsigned u = (unsigned)idx, a = u / (unsigned)nitems, b = a / (unsigned)cpn;
        it = (int)(u - a * (unsigned)nitems);
        cc = (int)(a - b * (unsigned)cpn);
        unsigned nn = b;
        if (full) {
            const unsigned c2 = b / (unsigned)G;
            grp_full = (int)(b - c2 * (unsigned)G);
            nn = c2;
        }
        n = nn;
    } else {
        it = (int)(idx % nitems);
        cc = (int)((idx / nitems) % cpn);
        n = idx / nitems / cpn;
        if (full) {
            grp_full = (int)(n % G);
            n /= G;
        }
    }
    int4 I = items[first + it];
    if (kind >= 0 && I.x != kind) return;
    const int npart = bands.npart;
    const i64 oh = h + 2 * pad, ow = W + 2 * pad;
    const int grp = full ? grp_full : psum - I.w;
    i64 c = (i64)grp * cpn + cc;
    if (I.x == 0) {
        int e = I.y, hr = I.z;
        int g = hr / (2 * pad), s = (hr / pad) % 2, r = hr % pad, x = e % W;
        int y = s == 0 ? r : pad + h + r;
        int pg = hband[hr];
        float *dst = buf + (((n * npart + g) * C + c) * oh + y) * ow + pad + x;
        const float *src = buf + (((n * npart + pg) * C + c) * oh + pad + hrow[hr]) * ow + pad;
        int q = hcol[e];
        float a = (q < 0) ? 0.f : src[q];
        int q1 = (q + 1 == bands.wl[pg]) ? 0 : q + 1;
        *dst = lerp2_ref(a, src[q1], htw[e]);
    } else {
        int g = I.y, y = I.z / pad, k = I.z % pad;
        float *rowp = buf + (((n * npart + g) * C + c) * oh + y) * ow;
        rowp[pad + bands.wl[g] + k] = rowp[pad + k];
    }
}

// ------------------------------------------------------------------------------------------------ masked conv
// entropy_conv2_data_to_col_gpu_v3{,_act}_batch (extension/entropy_conv_cuda_v2.cu:237-290, :326-379)
// FULL: every (cell, channel group) of the tensor in one launch, each scalar evaluated at its own step
// psum = group + row + col - the same chains and tree as the stepwise form, hence the same bits.
template <int GI, bool FULL>
__global__ void __launch_bounds__(256) ctx_conv_kernel(const float *__restrict__ in, const float *__restrict__ weight,
                                                       const float *__restrict__ bias, const float *__restrict__ act,
                                                       float *__restrict__ out, const int *__restrict__ order, int first,
                                                       int len, int nimg, i64 nscalars, int npart, int G, int go, int h,
                                                       int W, int pad_in, int pad_out, int constrain, int psum, Bands bands)
{
    const int lane = threadIdx.x & 31;
    const i64 wid = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= nscalars) return;
    int og, pn, tw, hp, tc;
    if (FULL) {
        // scalar id -> (col, global row, channel group, output-in-group, batch-image), col fastest
        i64 r = wid;
        tw = (int)(r % W); r /= W;
        hp = (int)(r % ((i64)h * npart)); r /= (i64)h * npart;
        tc = (int)(r % G); r /= G;
        og = (int)(r % go);
        pn = (int)(r / go);
        if (tw >= bands.wl[hp / h]) return;
        psum = tc + tw + hp;
    } else {
        // scalar id -> (cell k, output-in-group og, batch-image pn), cell fastest (the reference's blockIdx order)
        const int k = (int)(wid % len);
        og = (int)((wid / len) % go);
        pn = (int)(wid / len / go);
        const int hw = order[first + k];
        tw = hw % W; hp = hw / W;
        tc = psum - tw - hp;
    }
    const int b = pn / nimg;
    const int g = hp / h, th = hp % h;
    const int pout = tc * go + og;
    const int Ci = G * GI, Co = G * go;
    const i64 ih = h + 2 * pad_in, iw = W + 2 * pad_in;
    const i64 qn = (i64)pn * npart + g;
    const float *in_cell = in + (qn * Ci * ih + th + pad_in) * iw + tw + pad_in;
    const float *wgt = weight + ((i64)b * Co + pout) * Ci * 25;

    constexpr int NTH = 25 * GI;                 // live virtual lanes (25 or 75)
    constexpr int NV = (NTH + 31) / 32;          // virtual lanes per physical lane (1 or 3)
    float v[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NV; j++) {
        const int i = lane + 32 * j;
        if (i < NTH) {
            const int kw = i % 5, kh = (i / 5) % 5, m = i / 25;
            const int qh = hp - 2 + kh, pw = tw - 2 + kw;
            int nch = (constrain == 5 ? (psum - qh - pw) : (psum - qh - pw + 1)) * GI;
            if (nch > Ci) nch = Ci;
            const float *ip = in_cell + (i64)(kh - 2) * iw + (kw - 2);
            const float *wp = wgt + kh * 5 + kw;
            float acc = 0.f;
            for (int ti = m; ti < nch; ti += GI) acc = __fmaf_rn(ip[(i64)ti * ih * iw], wp[ti * 25], acc);
            v[j] = acc;
        }
    }
    // [t] += [t+64] (t < 64): virtual lanes l and l+32 absorb l+64 and l+96 (the latter never live)
    float s0 = __fadd_rn(v[0], v[2]);
    float s1 = __fadd_rn(v[1], 0.f);
    // [t] += [t+32] (t < 32)
    float sum = __fadd_rn(s0, s1);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum = __fadd_rn(sum, __shfl_down_sync(0xffffffffu, sum, off));
    if (lane == 0) {
        const i64 oh = h + 2 * pad_out, ow = W + 2 * pad_out;
        const int bidx = b * Co + pout;
        sum = __fadd_rn(sum, bias[bidx]);
        if (act != nullptr && sum < 0.f) sum = __fmul_rn(sum, act[bidx]);
        out[((qn * Co + pout) * oh + th + pad_out) * ow + tw + pad_out] = sum;
    }
}


// ------------------------------------------------------------------------------------------------ masked conv, tiled form
// Throughput form of the same arithmetic for the one-shot encoder.  A thread owns CPT cells (columns x, x+32, ...) of one
// row for one (net image, channel group tc) and produces their GO = 3 outputs.  It evaluates the reference's 128 virtual
// lanes one after the other - lane i = (m, kh, kw) is the chain acc = fma(in[k*GI+m], w[k*GI+m], acc) over the allowed groups
// k in ascending order - and folds them in the reference's tree ([t]+=[t+64], [t]+=[t+32], shuffle-down 16..1) depth first, so
// only a handful of partial sums are live.  Chain lengths depend on (tc, kh, kw) only: no divergence inside a block.  Each
// input value is loaded once (coalesced across the warp) for three FMAs; the weights of the block's (net, tc) sit in shared
// memory as one float4 per (channel, tap) and are read as warp-uniform 128-bit broadcasts.
#ifndef PCX_CT_CPT
#define PCX_CT_CPT 4
#endif
#ifndef PCX_CT_MINB
#define PCX_CT_MINB 4
#endif
constexpr int CT_CPT = PCX_CT_CPT;  // cells per thread
constexpr int CT_WARPS = 4;        // warps (= row tiles of 32*CPT columns) per block

struct CtAcc { float v[CT_CPT][3]; };

__device__ __forceinline__ CtAcc ct_zero()
{
    CtAcc a;
#pragma unroll
    for (int c = 0; c < CT_CPT; c++) a.v[c][0] = a.v[c][1] = a.v[c][2] = 0.f;
    return a;
}
__device__ __forceinline__ CtAcc ct_add(const CtAcc &a, const CtAcc &b)
{
    CtAcc r;
#pragma unroll
    for (int c = 0; c < CT_CPT; c++)
#pragma unroll
        for (int o = 0; o < 3; o++) r.v[c][o] = __fadd_rn(a.v[c][o], b.v[c][o]);
    return r;
}

struct CtCtx {
    const float *in_cell;       // input at (channel 0, row, column of cell 0), this lane
    const float4 *ws;           // shared weights [(ci * 25 + tap)] -> (og0, og1, og2, -)
    i64 chs;                    // channel stride of the padded input
    int iw, tc, G, c6;
};

// U = unroll factor of the channel-group loop, 0 = the compiler's choice.  The global-memory form needs the compiler's deep
// unrolling (16 groups of loads in flight hide the L1 / L2 latency) although it makes the kernel ~750 KB of straight-line
// code; the shared-memory form reads operands with ~30-cycle latency and is instead bound by instruction fetch at that
// size, so it keeps the 75 chain loops compact (U = 2: one iteration's loads overlap the previous one's FMAs).
template <int GI, int I, int U>
__device__ __forceinline__ CtAcc ct_chain(const CtCtx &x)
{
    CtAcc a = ct_zero();
    if (I < 25 * GI) {
        constexpr int kw = I % 5, kh = (I / 5) % 5, m = I / 25;
        int nk = x.tc + 4 - kh - kw + x.c6;
        nk = nk > x.G ? x.G : nk;
        const float *ip = x.in_cell + (i64)m * x.chs + (kh - 2) * x.iw + (kw - 2);
        const float4 *wp = x.ws + m * 25 + kh * 5 + kw;
        auto body = [&]() {
            const float4 w = *wp;
#pragma unroll
            for (int c = 0; c < CT_CPT; c++) {
                const float v = ip[32 * c];
                a.v[c][0] = __fmaf_rn(v, w.x, a.v[c][0]);
                a.v[c][1] = __fmaf_rn(v, w.y, a.v[c][1]);
                a.v[c][2] = __fmaf_rn(v, w.z, a.v[c][2]);
            }
            ip += (i64)GI * x.chs;
            wp += GI * 25;
        };
        if (U == 0) {
            for (int k = 0; k < nk; k++) body();
        } else {
#pragma unroll U
            for (int k = 0; k < nk; k++) body();
        }
    }
    return a;
}

// value of virtual lane T after the two shared-memory folds: ([T] + [T+64]) + ([T+32] + [T+96]); dead lanes hold +0.0f
template <int GI, int T, int U>
__device__ __forceinline__ CtAcc ct_leaf(const CtCtx &x)
{
    CtAcc lo = ct_add(ct_chain<GI, T, U>(x), ct_chain<GI, T + 64, U>(x));
    CtAcc hi = ct_add(ct_chain<GI, T + 32, U>(x), ct_zero());
    return ct_add(lo, hi);
}
// value of lane T after the shuffle-down steps 16 .. OFF
template <int GI, int T, int OFF, int U = 0>
struct CtTree {
    static __device__ __forceinline__ CtAcc run(const CtCtx &x)
    {
        CtAcc a = CtTree<GI, T, OFF * 2, U>::run(x);
        CtAcc b = CtTree<GI, T + OFF, OFF * 2, U>::run(x);
        return ct_add(a, b);
    }
};
template <int GI, int T, int U>
struct CtTree<GI, T, 32, U> {
    static __device__ __forceinline__ CtAcc run(const CtCtx &x) { return ct_leaf<GI, T, U>(x); }
};

// Runtime-loop form of the same tree for the shared-memory kernel: even with U = 1 the 75 template-expanded chain loops are
// 81 KB of code and the kernel stays instruction-fetch-bound (1.00 ms per layer at 2048x4096 against 2.53 ms for the fully
// unrolled 750 KB version - smaller was faster at every step).  Here the virtual lanes are walked by ONE loop in the tree's
// depth-first leaf order T = bitrev5(i); a leaf is ([T] + [T+64]) + ([T+32] + 0) as before, and the shuffle-down levels
// 16, 8, 4, 2, 1 are a binary-counter stack of five partial sums in named registers (level j combines when bit j of i is
// set: earlier + later, the operand order of CtTree).  Same chains, same adds, same order - a few hundred instructions.
template <int GI, int U>
__device__ __forceinline__ CtAcc ct_chain_rt(const CtCtx &x, int I)
{
    CtAcc a = ct_zero();
    if (I < 25 * GI) {
        const int m = I / 25, r = I - m * 25, kh = r / 5, kw = r - kh * 5;
        int nk = x.tc + 4 - kh - kw + x.c6;
        nk = nk > x.G ? x.G : nk;
        const int chs = (int)x.chs;
        const float *ip = x.in_cell + m * chs + (kh - 2) * x.iw + (kw - 2);
        const float4 *wp = x.ws + m * 25 + r;
#pragma unroll U
        for (int k = 0; k < nk; k++) {
            const float4 w = *wp;
#pragma unroll
            for (int c = 0; c < CT_CPT; c++) {
                const float v = ip[32 * c];
                a.v[c][0] = __fmaf_rn(v, w.x, a.v[c][0]);
                a.v[c][1] = __fmaf_rn(v, w.y, a.v[c][1]);
                a.v[c][2] = __fmaf_rn(v, w.z, a.v[c][2]);
            }
            ip += GI * chs;
            wp += GI * 25;
        }
    }
    return a;
}

template <int GI, int U>
__device__ __forceinline__ CtAcc ct_tree_rt(const CtCtx &x)
{
    CtAcc s0 = ct_zero(), s1 = ct_zero(), s2 = ct_zero(), s3 = ct_zero(), s4 = ct_zero(), acc = ct_zero();
#pragma unroll 1
    for (int i = 0; i < 32; i++) {
        const int T = ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
        const CtAcc lo = ct_add(ct_chain_rt<GI, U>(x, T), ct_chain_rt<GI, U>(x, T + 64));
        const CtAcc hi = ct_add(ct_chain_rt<GI, U>(x, T + 32), ct_zero());
        acc = ct_add(lo, hi);
        if (i & 1) {
            acc = ct_add(s0, acc);
            if (i & 2) {
                acc = ct_add(s1, acc);
                if (i & 4) {
                    acc = ct_add(s2, acc);
                    if (i & 8) {
                        acc = ct_add(s3, acc);
                        if (i & 16) acc = ct_add(s4, acc);
                        else s4 = acc;
                    } else s3 = acc;
                } else s2 = acc;
            } else s1 = acc;
        } else s0 = acc;
    }
    return acc;
}

template <int GI, int U = 0>
__global__ void __launch_bounds__(32 * CT_WARPS, PCX_CT_MINB) ctx_conv_tiled_kernel(const float *__restrict__ in, const float *__restrict__ weight,
                                                                       const float *__restrict__ bias, const float *__restrict__ act,
                                                                       const float *__restrict__ addsrc, float *__restrict__ out,
                                                                       int nimg, int npart, int G, int h, int W, int pad_in,
                                                                       int pad_out, int constrain, Bands bands, int step_lo, int step_hi)
{
    // step_lo / step_hi: only the scalars whose wavefront step tw + hp + tc lies in [step_lo, step_hi) are produced (slab-wise
    // one-shot encoding: the host codes the first slabs' symbols while the device computes the later ones)
    extern __shared__ float4 ct_ws[];
    const int tc = blockIdx.y, pn = blockIdx.z, b = pn / nimg;
    const int Ci = G * GI, Co = G * 3;
    // stage the weights of the three outputs of (net b, group tc): only the channel groups any chain can reach
    int gmax = tc + 4 + (constrain == 6 ? 1 : 0);
    gmax = gmax > G ? G : gmax;
    const int nw = gmax * GI * 25;
    for (int i = threadIdx.x; i < nw * 3; i += blockDim.x) {
        const int og = i / nw, r = i % nw;
        reinterpret_cast<float *>(ct_ws)[r * 4 + og] = weight[(((i64)b * Co + tc * 3 + og) * Ci) * 25 + r];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ntile = (W + 32 * CT_CPT - 1) / (32 * CT_CPT);
    const int rt = blockIdx.x * CT_WARPS + warp;
    const int Hf = h * npart;
    if (rt >= Hf * ntile) return;
    // consecutive warps (and blocks) take consecutive ROWS of one column tile: their 5-row windows overlap, so the block's
    // working set stays L1-resident
    const int hp = rt % Hf, x0 = (rt / Hf) * 32 * CT_CPT;
    const int g = hp / h, th = hp % h, wl = bands.wl[g];
    if (x0 >= wl) return;
    if (hp + tc + x0 >= step_hi || hp + tc + x0 + 32 * CT_CPT - 1 < step_lo) return;      // no cell of the tile in this slab
    const i64 ih = h + 2 * pad_in, iw = W + 2 * pad_in;
    const i64 qn = (i64)pn * npart + g;
    // W is a multiple of 32*CPT (launcher), so every column of the tile lies inside the padded row; columns >= wl are
    // computed on zero cells and discarded
    const int tw0 = x0 + lane;
    CtCtx x;
    x.chs = ih * iw;
    x.iw = (int)iw;
    x.tc = tc;
    x.G = G;
    x.c6 = constrain == 6 ? 1 : 0;
    x.ws = ct_ws;
    x.in_cell = in + (qn * Ci * ih + th + pad_in) * iw + tw0 + pad_in;
    CtAcc r = CtTree<GI, 0, 1, U>::run(x);
    const i64 oh = h + 2 * pad_out, ow = W + 2 * pad_out;
#pragma unroll
    for (int c = 0; c < CT_CPT; c++) {
        const int tw = tw0 + 32 * c;
        if (tw >= wl || tw + hp + tc < step_lo || tw + hp + tc >= step_hi) continue;
#pragma unroll
        for (int og = 0; og < 3; og++) {
            const int pout = tc * 3 + og, bidx = b * Co + pout;
            float sum = __fadd_rn(r.v[c][og], bias[bidx]);
            if (act != nullptr && sum < 0.f) sum = __fmul_rn(sum, act[bidx]);
            const i64 o = ((qn * Co + pout) * oh + th + pad_out) * ow + tw + pad_out;
            if (addsrc != nullptr) sum = __fadd_rn(sum, addsrc[o]);
            out[o] = sum;
        }
    }
}

// Shared-memory form of the tiled kernel (G * GI <= 42 input channels, i.e. valid_dim 56 = model-idx 3).  The kernel above
// re-reads every input value from L1 / L2 once per (tap, channel group it feeds): 12 % of the FP32 FMA rate, L1 hit rate 19 %,
// L2-latency-bound (profiles/r1v_ctx_conv_tiled_kernel.json).  Here a block owns CS_ROWS consecutive rows x 128 columns of one
// band plane, stages their whole 5x5 input window - (CS_ROWS + 4) rows x 132 columns x all input channels, 177 KB - ONCE, and
// loops over the G channel groups tc itself: each staged value feeds up to 25 taps x 3 outputs x G/2 groups from shared
// memory.  Eight warps: warp (r, par) takes row r and the channel groups tc = par, par + 2, ...; the two weight sets of an
// iteration are staged side by side.  Identical chains and fold tree (ct_chain / CtTree): bit-identical outputs.
constexpr int CS_ROWS = 4;
constexpr int CS_COLS = 32 * CT_CPT + 4;

template <int GI, int U>
__global__ void __launch_bounds__(64 * CS_ROWS, 1) ctx_conv_smem_kernel(const float *__restrict__ in, const float *__restrict__ weight,
                                                                       const float *__restrict__ bias, const float *__restrict__ act,
                                                                       const float *__restrict__ addsrc, float *__restrict__ out,
                                                                       int nimg, int npart, int G, int h, int W, int pad_in,
                                                                       int pad_out, int constrain, Bands bands, int step_lo, int step_hi,
                                                                       int tsplit)
{
    // tsplit: the channel-group pairs (tc, tc + 1) of a tile are dealt to `tsplit` blocks (pair index mod tsplit), so that a small
    // problem still fills the machine; a block stages only the input channels its highest group can reach.
    extern __shared__ float4 cs_smem[];
    const int Ci = G * GI, Co = G * 3;
    const int wslots = Ci * 25;                                            // float4 per weight set
    float4 *s_w = cs_smem;                                                 // [2][wslots]
    float *s_in = reinterpret_cast<float *>(cs_smem + 2 * wslots);         // [Ci][CS_ROWS + 4][CS_COLS]
    const int pn = blockIdx.z / tsplit, tg = blockIdx.z % tsplit, b = pn / nimg, g = blockIdx.y;
    if (2 * tg >= G) return;
    const int ntile = W / (32 * CT_CPT);
    const int x0 = (blockIdx.x % ntile) * 32 * CT_CPT, th0 = (blockIdx.x / ntile) * CS_ROWS;
    const int wl = bands.wl[g];
    if (x0 >= wl) return;
    // the block's scalars span the steps [g h + th0 + x0, g h + th0 + CS_ROWS - 1 + x0 + 127 + G - 1]: nothing to do outside the slab
    if (g * h + th0 + x0 >= step_hi || g * h + th0 + CS_ROWS - 1 + x0 + 32 * CT_CPT - 1 + G - 1 < step_lo) return;
    const int ih = h + 2 * pad_in, iw = W + 2 * pad_in;
    const i64 qn = (i64)pn * npart + g;
    constexpr int WR = CS_ROWS + 4;
    // ---- stage the input window: rows th0 - 2 .. th0 + CS_ROWS + 1, columns x0 - 2 .. x0 + 129 of the padded plane
    {
        const float *src = in + (qn * Ci * ih + th0 + pad_in - 2) * (i64)iw + x0 + pad_in - 2;
        // highest channel group of this block: the last pair it owns, tc_hi = t0_last + 1; its chains reach groups < tc_hi + 4 (+1)
        const int npair = (G + 1) / 2, last_pair = tg + ((npair - 1 - tg) / tsplit) * tsplit;
        int gtop = 2 * last_pair + 1 + 4 + (constrain == 6 ? 1 : 0);
        gtop = gtop > G ? G : gtop;
        const int per_ch = WR * CS_COLS, total = gtop * GI * per_ch;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int c = i / per_ch, r = (i - c * per_ch) / CS_COLS, col = i - c * per_ch - r * CS_COLS;
            if (th0 + pad_in - 2 + r < ih) cp_async4_ctx(s_in + i, src + ((i64)c * ih + r) * iw + col);
            else s_in[i] = 0.f;                                            // rows below the plane (h not a multiple of CS_ROWS)
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = warp % CS_ROWS, par = warp / CS_ROWS;
    const int th = th0 + r;
    const int c6 = constrain == 6 ? 1 : 0;
    CtCtx x;
    x.chs = WR * CS_COLS;
    x.iw = CS_COLS;
    x.G = G;
    x.c6 = c6;
    x.ws = s_w + par * wslots;
    x.in_cell = s_in + (r + 2) * CS_COLS + lane + 2;
    const i64 oh = h + 2 * pad_out, ow = W + 2 * pad_out;
    for (int t0 = 2 * tg; t0 < G; t0 += 2 * tsplit) {
        __syncthreads();                                                   // the previous pair's weight sets are no longer read
        // weight rows of the three outputs of (net b, group tc) for tc = t0, t0 + 1: only the reachable channel groups
        for (int q = 0; q < 2 && t0 + q < G; q++) {
            const int tcq = t0 + q;
            int gmax = tcq + 4 + c6;
            gmax = gmax > G ? G : gmax;
            const int nw = gmax * GI * 25;
            float *wd = reinterpret_cast<float *>(s_w + q * wslots);
            for (int i = threadIdx.x; i < nw * 3; i += blockDim.x) {
                const int og = i / nw, rr = i % nw;
                cp_async4_ctx(wd + rr * 4 + og, weight + (((i64)b * Co + tcq * 3 + og) * Ci) * 25 + rr);
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const int tc = t0 + par;
        if (tc >= G || th >= h) continue;
        const int st0 = g * h + th + tc + x0;              // step of the warp's first cell
        if (st0 >= step_hi || st0 + 32 * CT_CPT - 1 < step_lo) continue;
        x.tc = tc;
        CtAcc acc;                                                   // U = 8 / 9: runtime tree, chain loop x1 / x2
        if constexpr (U >= 8) acc = ct_tree_rt<GI, U - 7>(x);
        else acc = CtTree<GI, 0, 1, U>::run(x);
#pragma unroll
        for (int c = 0; c < CT_CPT; c++) {
            const int tw = x0 + lane + 32 * c;
            if (tw >= wl || st0 + lane + 32 * c < step_lo || st0 + lane + 32 * c >= step_hi) continue;
#pragma unroll
            for (int og = 0; og < 3; og++) {
                const int pout = tc * 3 + og, bidx = b * Co + pout;
                float sum = __fadd_rn(acc.v[c][og], bias[bidx]);
                if (act != nullptr && sum < 0.f) sum = __fmul_rn(sum, act[bidx]);
                const i64 o = ((qn * Co + pout) * oh + th + pad_out) * ow + tw + pad_out;
                if (addsrc != nullptr) sum = __fadd_rn(sum, addsrc[o]);
                out[o] = sum;
            }
        }
    }
}

// entropy_add_forward_kernel (extension/entropy_add_cuda.cu:25-44)
__global__ void ctx_add_kernel(float *__restrict__ y, const float *__restrict__ x, const int *__restrict__ order, int first,
                               int len, int nrep, int cpg, int npart, int C, int h, int W, int pad, int psum)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)len * cpg * nrep) return;
    int pn = (int)(idx % nrep);
    int k = (int)((idx / nrep) % len);
    int og = (int)(idx / nrep / len);
    int hw = order[first + k];
    int tw = hw % W, hp = hw / W, g = hp / h, th = hp % h;
    int tc = psum - tw - hp;
    i64 i = ((((i64)pn * npart + g) * C + (i64)tc * cpg + og) * (h + 2 * pad) + th + pad) * (W + 2 * pad) + tw + pad;
    y[i] = __fadd_rn(y[i], x[i]);
}

// d_input2_forward_kernel (extension/d_input_cuda_v2.cu:32-52)
__global__ void dinput_kernel(const float *__restrict__ sym, float *__restrict__ out, const int *__restrict__ order, int first,
                              int len, int nimg, int npart, int G, int h, int W, int pad, float bias, int rep,
                              i64 rep_stride, int psum)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)len * nimg) return;
    int k = (int)(idx % len);
    i64 n = idx / len;
    int hw = order[first + k];
    int tw = hw % W, hp = hw / W, g = hp / h, th = hp % h;
    int tc = psum - tw - hp;
    i64 i = (((n * npart + g) * G + tc) * (h + 2 * pad) + th + pad) * (W + 2 * pad) + tw + pad;
    float v = __fadd_rn(sym[idx], bias);
    for (int j = 0; j < rep; j++) out[i + j * rep_stride] = v;
}

// d_extract2_forward_kernel / d_extract2_batch_forward_kernel (extension/d_extract_cuda_v2.cu:34-52, :110-132)
__global__ void dextract_kernel(const float *__restrict__ in, float *__restrict__ out, const int *__restrict__ order, int first,
                                int len, int nrep, int npart, int G, int cpn, int h, int W, int psum, int nimg_batch,
                                i64 net_stride)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (i64)len * cpn * nrep) return;
    int ci = (int)(idx % cpn);
    int k = (int)((idx / cpn) % len);
    int n = (int)(idx / cpn / len);
    int hw = order[first + k];
    int tw = hw % W, hp = hw / W, g = hp / h, th = hp % h;
    int tc = psum - tw - hp;
    float v = in[((((i64)n * npart + g) * G * cpn + (i64)tc * cpn + ci) * h + th) * W + tw];
    if (nimg_batch > 0) out[(n / nimg_batch) * net_stride + (((i64)(n % nimg_batch) * len + k) * cpn + ci)] = v;
    else out[idx] = v;
}

}  // namespace
#include "pcx_ctx_step.cuh"
namespace {

// Fast path of the table operator for the codec's shape (3 Gaussians, 8 symbols): a block owns 128 consecutive rows; the
// three (n, 3) parameter arrays and the (n, 9) table pass through shared memory so that every global access is a coalesced
// run, and the per-row state lives in registers.
template <int FORM>
__global__ void __launch_bounds__(128) gmm_table38_kernel(float *__restrict__ logit, float *__restrict__ delta,
                                                          const float *__restrict__ mean, int n, float bias, float total,
                                                          float beta, float *__restrict__ cdf_f, int *__restrict__ cdf_i)
{
    __shared__ float sp[3][128 * 3];
    __shared__ float sc[128 * 9];
    const int r0 = blockIdx.x * 128, rows = min(128, n - r0), t = threadIdx.x;
    for (int i = t; i < rows * 3; i += 128) {
        sp[0][i] = logit[(i64)r0 * 3 + i];
        sp[1][i] = delta[(i64)r0 * 3 + i];
        sp[2][i] = mean[(i64)r0 * 3 + i];
    }
    __syncthreads();
    if (t < rows) {
        float w[3], d[3], mu[3], c[9];
#pragma unroll
        for (int i = 0; i < 3; i++) { w[i] = sp[0][t * 3 + i]; d[i] = sp[1][t * 3 + i]; mu[i] = sp[2][t * 3 + i]; }
        gmm_cdf_row<3, 8>(w, d, mu, 3, 8, bias, total, beta, FORM, c);
#pragma unroll
        for (int i = 0; i < 3; i++) { sp[0][t * 3 + i] = w[i]; sp[1][t * 3 + i] = d[i]; }
#pragma unroll
        for (int i = 0; i < 9; i++) sc[t * 9 + i] = c[i];
    }
    __syncthreads();
    for (int i = t; i < rows * 3; i += 128) {          // in place, like the reference
        logit[(i64)r0 * 3 + i] = sp[0][i];
        delta[(i64)r0 * 3 + i] = sp[1][i];
    }
    for (int i = t; i < rows * 9; i += 128) {
        if (cdf_f) cdf_f[(i64)r0 * 9 + i] = sc[i];
        if (cdf_i) cdf_i[(i64)r0 * 9 + i] = (int)sc[i];
    }
}

__global__ void gmm_table_kernel(float *__restrict__ logit, float *__restrict__ delta, const float *__restrict__ mean, int n,
                                 int ng, int nstep, float bias, float total, float beta, int form, float *__restrict__ cdf_f,
                                 int *__restrict__ cdf_i)
{
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    float w[PCX_MAX_GAUSS], d[PCX_MAX_GAUSS], mu[PCX_MAX_GAUSS];
    for (int i = 0; i < ng; i++) {
        w[i] = logit[(i64)r * ng + i];
        d[i] = delta[(i64)r * ng + i];
        mu[i] = mean[(i64)r * ng + i];
    }
    float c[33];
    gmm_cdf_row<0, 0>(w, d, mu, ng, nstep, bias, total, beta, form, c);
    for (int i = 0; i < ng; i++) {
        logit[(i64)r * ng + i] = w[i];                  // in place, like the reference
        delta[(i64)r * ng + i] = d[i];
    }
    for (int i = 0; i <= nstep; i++) {
        if (cdf_f) cdf_f[(i64)r * (nstep + 1) + i] = c[i];
        if (cdf_i) cdf_i[(i64)r * (nstep + 1) + i] = (int)c[i];
    }
}

// ------------------------------------------------------------------------------------------------ one-shot encoder kernels
// DInput2 over the whole tensor: every valid cell of every channel group gets symbol + bias, in the nb replicas.
__global__ void dinput_full_kernel(const float *__restrict__ sym, float *__restrict__ out, Bands bands, i64 total, int G, int h,
                                   int W, int pad, float bias, int rep, i64 rep_stride)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;                       // total = nimg * npart * G * h * W, layout of the symbol tensor
    i64 r = idx;
    int tw = (int)(r % W); r /= W;
    int th = (int)(r % h); r /= h;
    int tc = (int)(r % G);
    i64 n = r / G;                                  // image * npart + band
    if (tw >= bands.wl[(int)(n % bands.npart)]) return;
    i64 i = ((n * G + tc) * (h + 2 * pad) + th + pad) * (W + 2 * pad) + tw + pad;
    float v = __fadd_rn(sym[idx], bias);
    for (int j = 0; j < rep; j++) out[i + j * rep_stride] = v;
}

// EntropyAdd over the whole tensor (valid interior cells).
__global__ void ctx_add_full_kernel(float *__restrict__ y, const float *__restrict__ x, Bands bands, i64 total, int C, int h, int W,
                                    int pad)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;                       // total = planes * C * h * W
    i64 r = idx;
    int tw = (int)(r % W); r /= W;
    int th = (int)(r % h); r /= h;
    i64 pc = r;                                     // plane * C + channel
    i64 plane = pc / C;
    if (tw >= bands.wl[(int)(plane % bands.npart)]) return;
    i64 i = (pc * (h + 2 * pad) + th + pad) * (W + 2 * pad) + tw + pad;
    y[i] = __fadd_rn(y[i], x[i]);
}

// DExtract2Batch + EntropyBatchGmmTable + DExtract2(label) for a range of wavefront steps, rows written in CODING order:
// row = image * rows_per_image + (rowbase[s] - rowbase[s0]) + k, where k runs over the cell window of step s
// (cell = order[wfirst[s] + k], channel group = s - row - col) - the order in which the stepwise loop feeds the coder.
__global__ void gmm_ordered_kernel(const float *__restrict__ params, const float *__restrict__ data, const int *__restrict__ order,
                                   const int *__restrict__ wfirst, const int *__restrict__ rowbase, int s0, int s1, int nimg,
                                   int npart, int G, int go, int h, int W, int ng, int nstep, float bias, float total, float beta,
                                   int *__restrict__ cdf_i, int *__restrict__ lab_i)
{
    const int per = rowbase[s1] - rowbase[s0];
    const i64 t = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (i64)per * nimg) return;
    const int img = (int)(t / per);
    const int j = (int)(t % per) + rowbase[s0];
    int lo = s0, hi = s1;                          // rowbase[lo] <= j < rowbase[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (rowbase[mid] <= j) lo = mid; else hi = mid;
    }
    const int s = lo, k = j - rowbase[lo];
    const int hw = order[wfirst[s] + k];
    const int tw = hw % W, hp = hw / W, g = hp / h, th = hp % h;
    const int tc = s - tw - hp;
    const int Co = G * go;
    // ng == 3, nstep == 8 (checked by the launcher): the row state stays in registers
    float w[3], d[3], mu[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const i64 cell = ((i64)tc * go + i) * h * W + (i64)th * W + tw;
        w[i] = params[(((i64)(0 * nimg + img) * npart + g) * Co) * h * W + cell];
        d[i] = params[(((i64)(1 * nimg + img) * npart + g) * Co) * h * W + cell];
        mu[i] = params[(((i64)(2 * nimg + img) * npart + g) * Co) * h * W + cell];
    }
    float c[9];
    gmm_cdf_row<3, 8>(w, d, mu, 3, 8, bias, total, beta, 0, c);
#pragma unroll
    for (int i = 0; i <= 8; i++) cdf_i[t * 9 + i] = (int)c[i];
    lab_i[t] = (int)data[((((i64)img * npart + g) * G + tc) * h + th) * W + tw];
}

// entropy_gmm_forward_kernel (extension/entropy_gmm_cuda.cu:36-69), loss only
__global__ void gmm_nll_kernel(const float *__restrict__ bottom_weight, const float *__restrict__ bottom_delta,
                               const float *__restrict__ bottom_mean, const float *__restrict__ label, float *__restrict__ loss,
                               int n, int ng)
{
    int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= n) return;
    float s2 = 1. / sqrt(float(2.0));
    float sum_p = 0;
    for (int i = 0; i < ng; i++) {
        float xa = label[index] - 0.5 - bottom_mean[index * ng + i];
        float xb = label[index] + 0.5 - bottom_mean[index * ng + i];
        float id = 1. / bottom_delta[index * ng + i];
        float fa = 0.5 + 0.5 * erf(xa * id * s2);
        float fb = 0.5 + 0.5 * erf(xb * id * s2);
        float p = fb - fa;
        sum_p = sum_p + bottom_weight[index * ng + i] * p;
    }
    loss[index] = -log(sum_p + 0.0000001);
}

// ------------------------------------------------------------------------------------------------ fused wavefront step
// One cooperative launch per wavefront step (decoder): DInput2 of the symbols decoded at the previous step, the 12 masked
// layers (+ residual adds) at the cells of the current plane window, DExtract2Batch + GMM table - separated by grid-wide
// barriers instead of kernel boundaries (1 launch per step instead of 32).  The halo of every layer input is never
// materialised: a tap that falls on a halo / right-wrap cell evaluates the causal 2-tap interpolation (lerp2_ref, same
// expression as ctx_pad_kernel) from the interior cells it is a function of.  Those cells lie on earlier-or-equal planes,
// i.e. they are final when the tap is allowed by the mask - so every scalar sees the same input values as in
// the stepwise operator path and runs the same chains and fold tree: bit-identical CDFs.  Activations written during the
// launch are read with ld.global.cg (L2), never through L1.
// in / out / add are the engine's own CHANNELS-LAST scratch [plane][row][col][cp]: the 16*gi input channels of a cell are one
// 64*gi-byte run, so a lane that owns a filter tap fetches all its channel groups with a few 128-bit loads.
// start: device copy of the plane prefix of `order`; planes [p0, p0 + np) form the window of this step, cut into nchunk runs
// (step_chunk_of); block i works on runs i, i + gridDim.x, ...  For every (layer, run) item the three weight rows are staged
// in one of two shared-memory buffers while the previous item is being computed.
__global__ void __launch_bounds__(STEP_THREADS) wave_step_kernel(const __grid_constant__ StepNet d, const int *__restrict__ start,
                                                                 int step, int p0, int np, int S, int nchunk, int pfirst, int pcount,
                                                                 unsigned bar_base)
{
    extern __shared__ __align__(16) float step_ws[];  // 2 buffers of 4 * wstride weights + 8 (bias, slope)
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int h = d.h, W = d.W, pad = d.pad, G = d.G, nrep = d.nb * d.nimg;
    const int wstride = G * 3 * 25;
    unsigned target = bar_base;
    step_stamp(d, step, 0);
    const int first = start[p0], count = start[p0 + np] - start[p0];
    const int nmy = blockIdx.x < nchunk ? (nchunk - 1 - blockIdx.x) / gridDim.x + 1 : 0;      // my runs per layer
    const int nitems = nmy * d.nlayers;
    // the block's runs are the same for every layer: decode them ONCE (the walk over the planes costs three integer
    // divisions per plane - done per item by every thread it was half of the kernel's instructions)
    __shared__ StepChunk s_chunk[STEP_MAX_RUNS];
    __shared__ int s_cache_base[STEP_MAX_RUNS + 1];       // position of a run's first cell in the block's work list
    __shared__ StepCellRec s_cell[STEP_CACHE_CELLS];
    __shared__ StepTapOff s_tap[STEP_CACHE_CELLS * 25];
    if (threadIdx.x < nmy && threadIdx.x < STEP_MAX_RUNS)
        s_chunk[threadIdx.x] = step_chunk_of(start, blockIdx.x + threadIdx.x * gridDim.x, p0, np, S, d.nb, d.nimg);
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < nmy && i < STEP_MAX_RUNS; i++) { s_cache_base[i] = acc; acc += s_chunk[i].ncell; }
        s_cache_base[nmy < STEP_MAX_RUNS ? nmy : STEP_MAX_RUNS] = acc;
    }
    __syncthreads();
    {
        // resolve the (cell, tap) geometry of the first STEP_CACHE_CELLS cells once for all layers: warp per cell, lane per tap
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
        const int nruns = nmy < STEP_MAX_RUNS ? nmy : STEP_MAX_RUNS;
        const int ncached = s_cache_base[nruns] < STEP_CACHE_CELLS ? s_cache_base[nruns] : STEP_CACHE_CELLS;
        for (int f = warp; f < ncached; f += nwarp) {
            int ci = 0;
            while (ci + 1 < nruns && s_cache_base[ci + 1] <= f) ci++;
            const StepChunk ch = s_chunk[ci];
            const int k = f - s_cache_base[ci];
            const int pfirst_ = start[ch.plane], pcells = start[ch.plane + 1] - pfirst_;
            int img = ch.img0, ce = ch.rem0 + k;
            while (ce >= pcells) { ce -= pcells; img++; }
            const int pn = ch.net * d.nimg + img;
            const int4 cinfo = d.cell[pfirst_ + ce];
            if (lane == 0) s_cell[f] = StepCellRec{pn, cinfo.z, cinfo.w, cinfo.x};
            if (lane < 25) {
                const int kw = lane % 5, kh = lane / 5;
                s_tap[f * 25 + lane] = step_resolve_off(d, pn, cinfo.z, cinfo.w + kh - 2, cinfo.x + kw - 2);
            }
        }
    }                                                     // made visible by the grid barrier (block-level __syncthreads inside)
    if (nitems > 0) {
        const StepChunk c0 = s_chunk[0];
        step_stage_weights(d, d.L[0], c0.net, step - c0.plane, step_ws, wstride);
    }
    cp_async_commit();
    // ---- DInput2: the symbols decoded at step - 1 enter the padded input (3 replicas)
    if (pcount > 0) {
        const i64 ih = h + 2 * pad, iw = W + 2 * pad;
        const int cp0 = d.L[0].cp_in;
        const i64 rep_stride = (i64)d.nimg * d.npart * ih * iw * cp0;
        float *out = const_cast<float *>(d.L[0].in);
        for (int t = gtid; t < pcount * d.nimg; t += nthr) {
            const int k = t % pcount, n = t / pcount;
            const int4 ci = d.cell[pfirst + k];
            const int tw = ci.x, hp = ci.y, g = ci.z, th = ci.w;
            const int tc = step - 1 - tw - hp;
            const float v = __fadd_rn(d.prev[t], d.input_bias);
            const i64 i = ((((i64)n * d.npart + g) * ih + th + pad) * iw + tw + pad) * cp0 + tc;
            for (int j = 0; j < d.nb; j++) out[i + j * rep_stride] = v;
            d.sym_nchw[((((i64)n * d.npart + g) * G + tc) * ih + th + pad) * iw + tw + pad] = v;
        }
    }
    target += gridDim.x;
    grid_barrier(d.bar, target);
    step_stamp(d, step, 1);
    // ---- the masked layers
    int it = 0;
    for (int L = 0; L < d.nlayers; L++) {
        for (int ci = 0; ci < nmy; ci++, it++) {
            const StepChunk ch = ci < STEP_MAX_RUNS ? s_chunk[ci] : step_chunk_of(start, blockIdx.x + ci * gridDim.x, p0, np, S, d.nb, d.nimg);
            if (ci > 0) __syncthreads();               // every warp is done with item it-1: its buffer may be refilled
            if (it + 1 < nitems) {
                const int nci = ci + 1 < nmy ? ci + 1 : 0, nL = ci + 1 < nmy ? L : L + 1;
                const StepChunk nc = nci < STEP_MAX_RUNS ? s_chunk[nci] : step_chunk_of(start, blockIdx.x + nci * gridDim.x, p0, np, S, d.nb, d.nimg);
                step_stage_weights(d, d.L[nL], nc.net, step - nc.plane, step_ws + ((it + 1) & 1) * (4 * wstride + 8), wstride);
            }
            cp_async_commit();
            cp_async_wait<1>();                        // this item's rows have landed; the next item's may still be in flight
            __syncthreads();
            const float *ws = step_ws + (it & 1) * (4 * wstride + 8);
            const int cbase = ci < STEP_MAX_RUNS ? s_cache_base[ci] : STEP_CACHE_CELLS;
            if (d.L[L].gi == 1) step_conv_phase<1>(d, d.L[L], step, ch, start, ws, wstride, s_tap, s_cell, cbase);
            else step_conv_phase<3>(d, d.L[L], step, ch, start, ws, wstride, s_tap, s_cell, cbase);
        }
        target += gridDim.x;
        grid_barrier(d.bar, target);
        step_stamp(d, step, 2 + L);
    }
    cp_async_wait<0>();
    // ---- DExtract2Batch + GMM table, rows (image, cell of the window)
    const StepLayer &last = d.L[d.nlayers - 1];
    for (int t = gtid; t < count * d.nimg; t += nthr) {
        const int k = t % count, img = t / count;
        const int4 ci = d.cell[first + k];
        const int tw = ci.x, hp = ci.y, g = ci.z, th = ci.w;
        const int tc = step - tw - hp;
        float w[3], dl[3], mu[3];                          // ng == 3, nstep == 8 (checked by the launcher)
#pragma unroll
        for (int i = 0; i < 3; i++) {                     // 3 nets x 3 components, all loads in flight together
            const float *pp = last.out + (((i64)img * d.npart + g) * h + th) * W * last.cp_out + (i64)tw * last.cp_out + tc * 3 + i;
            const i64 net_stride = (i64)d.nimg * d.npart * h * W * last.cp_out;
            w[i] = __ldcg(pp);
            dl[i] = __ldcg(pp + net_stride);
            mu[i] = __ldcg(pp + 2 * net_stride);
        }
        float c[9];
        gmm_cdf_row<3, 8>(w, dl, mu, 3, 8, d.gmm_bias, d.gmm_total, d.gmm_beta, 0, c);
#pragma unroll
        for (int i = 0; i <= 8; i++) d.cdf[(i64)t * 9 + i] = (int)c[i];
    }
    step_stamp(d, step, 2 + d.nlayers);
}

}  // namespace

extern "C" {

int pcx_ctx_pad_step(float *d_buf, int nrep, int npart, int G, int cpn, int h, int W, int pad, int psum, const int *wl,
                     const int *d_band, const int *d_row, const int *d_col, const float *d_tw, const int *d_items,
                     const int *h_pstart, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_buf && d_band && d_row && d_col && d_tw && d_items && h_pstart, "null pointer");
    PCX_REQUIRE(nrep > 0 && G > 0 && cpn > 0 && h > 0 && W > 0 && pad > 0, "bad shape");
    const int Hf = h * npart;
    if (psum < 0 || psum >= Hf + W + pad + G - 2) return PCX_OK;          // :98
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W + pad - 2 ? psum + 1 : Hf + W + pad - 1;
    int nitems = h_pstart[en] - h_pstart[st];
    if (nitems <= 0) return PCX_OK;
    i64 total = (i64)nitems * cpn * nrep;
    ctx_pad_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(d_buf, b, d_band, d_row, d_col, d_tw,
                                                                           (const int4 *)d_items, h_pstart[st], nitems, nrep,
                                                                           cpn, G * cpn, h, W, pad, psum, 0, -1);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_ctx_conv_step(const float *d_in, const float *d_weight, const float *d_bias, const float *d_act, float *d_out, int nb,
                      int nimg, int npart, int G, int gi, int go, int h, int W, int pad_in, int pad_out, int constrain,
                      int psum, const int *d_order, const int *h_start, void *stream)
{
    PCX_REQUIRE(d_in && d_weight && d_bias && d_out && d_order && h_start, "null pointer");
    PCX_REQUIRE(nb > 0 && nimg > 0 && npart > 0 && G > 0 && go > 0 && h > 0 && W > 0, "bad shape");
    PCX_REQUIRE(gi == 1 || gi == 3, "input channels per group must be 1 or 3 (got %d)", gi);
    PCX_REQUIRE(constrain == 5 || constrain == 6, "constrain must be 5 or 6");
    const int Hf = h * npart;
    if (psum < 0 || psum >= Hf + W + G - 2) return PCX_OK;               // psum < mod_ (:398)
    Window w = wave_window(h_start, psum, G, Hf, W);
    if (w.count <= 0) return PCX_OK;
    i64 nscalars = (i64)nb * nimg * go * w.count;
    PCX_REQUIRE(nscalars < (1ll << 26), "wavefront too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int threads = 256;
    int blocks = grid_for(nscalars * 32, threads);
    Bands nob = {};
    if (gi == 1)
        ctx_conv_kernel<1, false><<<blocks, threads, 0, s>>>(d_in, d_weight, d_bias, d_act, d_out, d_order, w.first, w.count, nimg,
                                                             nscalars, npart, G, go, h, W, pad_in, pad_out, constrain, psum, nob);
    else
        ctx_conv_kernel<3, false><<<blocks, threads, 0, s>>>(d_in, d_weight, d_bias, d_act, d_out, d_order, w.first, w.count, nimg,
                                                             nscalars, npart, G, go, h, W, pad_in, pad_out, constrain, psum, nob);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_ctx_add_step(float *d_y, const float *d_x, int nrep, int npart, int G, int cpg, int h, int W, int pad, int psum,
                     const int *d_order, const int *h_start, void *stream)
{
    PCX_REQUIRE(d_y && d_x && d_order && h_start, "null pointer");
    const int Hf = h * npart;
    if (psum < 0 || psum > Hf + W + G - 2) return PCX_OK;                // psum <= mod_ (:56)
    Window w = wave_window(h_start, psum, G, Hf, W);
    i64 total = (i64)w.count * cpg * nrep;
    if (total <= 0) return PCX_OK;
    ctx_add_kernel<<<grid_for(total, 128), 128, 0, (cudaStream_t)stream>>>(d_y, d_x, d_order, w.first, w.count, nrep, cpg, npart,
                                                                           G * cpg, h, W, pad, psum);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_dinput_step(const float *d_sym, float *d_out, int nimg, int npart, int G, int h, int W, int pad, float bias, int rep,
                    int psum, const int *d_order, const int *h_start, void *stream)
{
    PCX_REQUIRE(d_sym && d_out && d_order && h_start, "null pointer");
    const int Hf = h * npart;
    i64 rep_stride = (i64)nimg * npart * G * (h + 2 * pad) * (W + 2 * pad);
    cudaStream_t s = (cudaStream_t)stream;
    if (psum == 0) {                                                       // :68-70
        PCX_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * rep * rep_stride, s));
        return PCX_OK;
    }
    if (psum < 0 || psum > Hf + W + G - 2) return PCX_OK;
    psum -= 1;
    Window w = wave_window(h_start, psum, G, Hf, W);
    i64 total = (i64)w.count * nimg;
    if (total <= 0) return PCX_OK;
    dinput_kernel<<<grid_for(total, 128), 128, 0, s>>>(d_sym, d_out, d_order, w.first, w.count, nimg, npart, G, h, W, pad, bias,
                                                       rep, rep_stride, psum);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_dextract_step(const float *d_in, float *d_out, int nrep, int npart, int G, int cpn, int h, int W, int psum, int batch,
                      int lag, const int *d_order, const int *h_start, int *h_count, void *stream)
{
    PCX_REQUIRE(d_in && d_out && d_order && h_start && h_count, "null pointer");
    PCX_REQUIRE(!batch || nrep % 3 == 0, "batch extraction needs 3 nets");
    const int Hf = h * npart;
    const int mod = Hf + W + G - 2;
    cudaStream_t s = (cudaStream_t)stream;
    if (lag) {                                                             // label_ == false branch (:85-98)
        if (psum == 0) {
            PCX_CUDA(cudaMemsetAsync(d_out, 0, sizeof(float) * (size_t)nrep * cpn * Hf * W, s));
            return PCX_OK;
        }
        if (psum > mod) return PCX_OK;
        psum -= 1;
    } else if (psum >= mod) {
        return PCX_OK;
    }
    Window w = wave_window(h_start, psum, G, Hf, W);
    int nimg = batch ? nrep / 3 : nrep;
    *h_count = nimg * w.count;
    i64 total = (i64)w.count * cpn * nrep;
    if (total <= 0) return PCX_OK;
    dextract_kernel<<<grid_for(total, 128), 128, 0, s>>>(d_in, d_out, d_order, w.first, w.count, nrep, npart, G, cpn, h, W, psum,
                                                         batch ? nimg : 0, (i64)cpn * Hf * W * nimg);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_gmm_table(float *d_logit, float *d_delta, const float *d_mean, int n, int ng, int nstep, float bias, float total,
                  float beta, int form, float *d_cdf_f, int *d_cdf_i, void *stream)
{
    PCX_REQUIRE(d_logit && d_delta && d_mean && (d_cdf_f || d_cdf_i), "null pointer");
    PCX_REQUIRE(ng >= 1 && ng <= PCX_MAX_GAUSS, "num_gaussian %d > 16 (entropy_gmm_table_cuda.cu:13)", ng);
    PCX_REQUIRE(nstep >= 2 && nstep <= 32, "nstep %d out of range", nstep);
    PCX_REQUIRE(form == 0 || form == 1, "form %d", form);
    if (n <= 0) return PCX_OK;                                             // tn > 0 (:163)
    if (ng == 3 && nstep == 8) {
        if (form == 0)
            gmm_table38_kernel<0><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_logit, d_delta, d_mean, n, bias, total, beta, d_cdf_f, d_cdf_i);
        else
            gmm_table38_kernel<1><<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_logit, d_delta, d_mean, n, bias, total, beta, d_cdf_f, d_cdf_i);
    } else {
        gmm_table_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_logit, d_delta, d_mean, n, ng, nstep, bias, total,
                                                                             beta, form, d_cdf_f, d_cdf_i);
    }
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_gmm_nll(const float *d_w, const float *d_delta, const float *d_mean, const float *d_label, float *d_loss, int n,
                int ng, void *stream)
{
    PCX_REQUIRE(d_w && d_delta && d_mean && d_label && d_loss, "null pointer");
    PCX_REQUIRE(ng >= 1 && ng <= PCX_MAX_GAUSS, "num_gaussian %d out of range", ng);
    if (n <= 0) return PCX_OK;
    gmm_nll_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(d_w, d_delta, d_mean, d_label, d_loss, n, ng);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ wavefront engine
// The whole serial loop of EntEncoder / EntDecoder (pseudo_codec.py:97-114, :145-160) in native code: per step the same
// launch sequence as the operator-by-operator path (DInput2 -> 12 x [EntropyCtxPadRun2 -> EntropyConv2Batch (+ EntropyAdd)]
// -> DExtract2Batch -> EntropyBatchGmmTable), through the SAME kernels - so every CDF entry is bit-identical - but
// without a Python / ctypes round trip per operator, with int32 CDF rows copied to pinned host memory (only the live
// rows) and the host range coder called from the same loop.  The encoder is software-pipelined: the kernels of step
// s+1 are enqueued before the host codes the tables of step s (labels come from the device-resident symbol tensor, so
// there is no host dependency); the decoder has a true dependency per step (the symbols decoded at step s feed step s+1).
namespace {

struct PinnedPool {
    int32_t *cdf[2] = {nullptr, nullptr};
    float *lab[2] = {nullptr, nullptr};
    float *d_prev = nullptr;
    size_t rows = 0;
};
// Everything the engines keep between calls lives in one record PER DEVICE (the reference API allows several devices in one
// process: every op carries its `device`, pseudo_codec has --gpu-id), guarded by a mutex held for the duration of an
// encode / decode call - two host threads on one device take turns, two devices do not touch each other's buffers.
struct WaveDev {
    std::mutex mu;
    PinnedPool pin;
    unsigned *bar = nullptr;                 // grid barrier counter of the step kernel
    float *step_scratch = nullptr;           // channels-last scratch of the step kernel (+ cell table)
    size_t step_scratch_bytes = 0;
    int *d_start = nullptr;                  // device copy of the plane prefix
    size_t d_start_cap = 0;
    cudaStream_t s_copy = nullptr;           // one-shot encoder: CDF rows leave on their own stream
    cudaEvent_t slab_ev[16] = {};
};
std::mutex g_wave_devs_mu;
WaveDev *g_wave_devs[64] = {nullptr};

WaveDev *wave_dev()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lk(g_wave_devs_mu);
    if (!g_wave_devs[dev]) g_wave_devs[dev] = new WaveDev();
    return g_wave_devs[dev];
}


// Per-image host coding in parallel: the nimg bitstreams are independent.  T = pcx_host_coder_threads(nimg) threads (the caller
// is thread 0) share the images round-robin, so a rank never runs more coder threads than its share of the host cores - eight
// ranks x eight images used to be 64 spinning threads on a 32-core box and decode throughput collapsed at N = 8.  Workers
// spin briefly on a generation counter (a step's coding job is tens of microseconds, below a condition variable's wake-up
// latency), then block on a condition variable; the caller waits the same way.  The pool lives for one encode / decode.
struct CoderPool {
    pcx_coder *const *coders = nullptr;
    int nimg = 1, ncode = 8, nthreads = 1;
    bool encode = true;
    // job of the current generation
    const int32_t *cdf = nullptr;
    int32_t *lab_i = nullptr;     // encode: symbols in
    float *lab_f = nullptr;       // decode: symbols out
    int per = 0;
    std::atomic<int> gen{0}, done{0}, status{0};
    std::atomic<bool> quit{false};
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> threads;

    void work(int t)
    {
        for (int im = t; im < nimg; im += nthreads) {
            int rc;
            if (encode) rc = pcx_coder_encodes(coders[im], cdf + (size_t)im * per * (ncode + 1), ncode, lab_i + (size_t)im * per, per);
            else rc = pcx_coder_decodes(coders[im], cdf + (size_t)im * per * (ncode + 1), ncode, per, lab_f + (size_t)im * per);
            if (rc < 0) status.store(rc);
        }
    }
    void loop(int t)
    {
        int seen = 0;
        for (;;) {
            int spins = 0;
            while (gen.load(std::memory_order_acquire) == seen && !quit.load(std::memory_order_relaxed)) {
                if (++spins < 4000) { __builtin_ia32_pause(); continue; }
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return gen.load(std::memory_order_acquire) != seen || quit.load(); });
            }
            if (quit.load()) return;
            seen = gen.load(std::memory_order_acquire);
            work(t);
            if (done.fetch_add(1, std::memory_order_acq_rel) + 1 == nthreads - 1) {
                std::lock_guard<std::mutex> lk(mu);
                cv_done.notify_one();
            }
        }
    }
    void start(pcx_coder *const *c, int n, int nc, bool enc)
    {
        coders = c; nimg = n; ncode = nc; encode = enc;
        nthreads = pcx_host_coder_threads(nimg);
        for (int t = 1; t < nthreads; t++) threads.emplace_back([this, t] { loop(t); });
    }
    int run(const int32_t *cdf_, int32_t *li, float *lf, int per_)
    {
        cdf = cdf_; lab_i = li; lab_f = lf; per = per_;
        done.store(0, std::memory_order_relaxed);
        {
            std::lock_guard<std::mutex> lk(mu);
            gen.fetch_add(1, std::memory_order_release);
        }
        cv_work.notify_all();
        work(0);
        int spins = 0;
        while (done.load(std::memory_order_acquire) < nthreads - 1) {
            if (++spins < 4000) { __builtin_ia32_pause(); continue; }
            std::unique_lock<std::mutex> lk(mu);
            cv_done.wait(lk, [&] { return done.load(std::memory_order_acquire) >= nthreads - 1; });
        }
        return status.load();
    }
    ~CoderPool()
    {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit.store(true);
        }
        cv_work.notify_all();
        for (auto &t : threads) t.join();
    }
};

int ensure_pinned(WaveDev &D, size_t rows, int nstep)
{
    if (rows <= D.pin.rows) return PCX_OK;
    for (int i = 0; i < 2; i++) {
        if (D.pin.cdf[i]) cudaFreeHost(D.pin.cdf[i]);
        if (D.pin.lab[i]) cudaFreeHost(D.pin.lab[i]);
        // mapped: the fused decoder step writes CDF rows / reads symbols through these buffers directly (zero-copy)
        PCX_CUDA(cudaHostAlloc((void **)&D.pin.cdf[i], rows * (nstep + 1) * sizeof(int32_t), cudaHostAllocMapped | cudaHostAllocPortable));
        PCX_CUDA(cudaHostAlloc((void **)&D.pin.lab[i], rows * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
    }
    D.pin.rows = rows;
    return PCX_OK;
}

// one wavefront step on the device: returns the number of symbols (rows of the CDF table) it produced
int wave_launch_step(const pcx_wave_net &n, int step, const float *d_prev, int *count, cudaStream_t s)
{
    const int nrep = n.nb * n.nimg;
    int rc = pcx_dinput_step(d_prev, n.layers[0].in, n.nimg, n.npart, n.G, n.h, n.W, n.pad, n.input_bias, n.nb, step, n.d_order,
                             n.h_start, s);
    if (rc < 0) return rc;
    for (int L = 0; L < n.nlayers; L++) {
        const pcx_wave_layer &l = n.layers[L];
        const i64 out_elems = (i64)nrep * n.npart * n.G * l.go * (n.h + 2 * l.pad_out) * (n.W + 2 * l.pad_out);
        if (step == 0) PCX_CUDA(cudaMemsetAsync(l.out, 0, sizeof(float) * out_elems, s));       // entropy_conv_cuda_v2.cu:307-309
        rc = pcx_ctx_pad_step(l.in, nrep, n.npart, n.G, l.gi, n.h, n.W, n.pad, l.input_layer ? step - 1 : step, n.wl, n.d_band,
                              n.d_row, n.d_col, n.d_tw, n.d_items, n.h_pstart, s);
        if (rc < 0) return rc;
        rc = pcx_ctx_conv_step(l.in, l.weight, l.bias, l.act, l.out, n.nb, n.nimg, n.npart, n.G, l.gi, l.go, n.h, n.W, n.pad, l.pad_out,
                               l.constrain, step, n.d_order, n.h_start, s);
        if (rc < 0) return rc;
        if (l.add) {
            rc = pcx_ctx_add_step(l.out, l.add, nrep, n.npart, n.G, l.go, n.h, n.W, l.pad_out, step, n.d_order, n.h_start, s);
            if (rc < 0) return rc;
        }
    }
    const pcx_wave_layer &last = n.layers[n.nlayers - 1];
    *count = 0;
    rc = pcx_dextract_step(last.out, n.d_params, nrep, n.npart, n.G, last.go, n.h, n.W, step, 1, 0, n.d_order, n.h_start, count, s);
    if (rc < 0) return rc;
    const i64 stride = (i64)last.go * n.h * n.npart * n.W * n.nimg;
    return pcx_gmm_table(n.d_params, n.d_params + stride, n.d_params + 2 * stride, *count, n.ng, n.nstep, n.gmm_bias, n.gmm_total,
                         n.gmm_beta, 0, nullptr, n.d_cdf, s);
}

int wave_check(const pcx_wave_net *net)
{
    PCX_REQUIRE(net != nullptr, "null net");
    const pcx_wave_net &n = *net;
    PCX_REQUIRE(n.nlayers >= 1 && n.nlayers <= PCX_WAVE_MAX_LAYERS, "nlayers %d", n.nlayers);
    PCX_REQUIRE(n.nb == 3 && n.nimg >= 1 && n.npart >= 1 && n.npart <= PCX_MAX_PART && n.G >= 1 && n.h >= 1 && n.W >= 1 && n.pad >= 1, "bad net shape");
    PCX_REQUIRE(n.wl && n.d_band && n.d_row && n.d_col && n.d_tw && n.d_items && n.h_pstart && n.d_order && n.h_start && n.d_params && n.d_cdf && n.d_prev, "null table / buffer");
    for (int L = 0; L < n.nlayers; L++)
        PCX_REQUIRE(n.layers[L].weight && n.layers[L].bias && n.layers[L].in && n.layers[L].out, "layer %d has a null pointer", L);
    PCX_REQUIRE(n.layers[n.nlayers - 1].pad_out == 0, "the last layer must have pad_out 0");
    return PCX_OK;
}

}  // namespace

extern "C" {

int pcx_wave_steps(const pcx_wave_net *net) { return net ? net->h * net->npart + net->W + net->G - 2 : PCX_EINVAL; }

int pcx_wave_encode(const pcx_wave_net *net, const float *d_data, pcx_coder *const *coders, long long *n_symbols, void *stream)
{
    int rc = wave_check(net);
    if (rc < 0) return rc;
    PCX_REQUIRE(d_data && coders, "null data / coders");
    for (int i = 0; i < net->nimg; i++) PCX_REQUIRE(coders[i] != nullptr, "null coder for image %d", i);
    const pcx_wave_net &n = *net;
    cudaStream_t s = (cudaStream_t)stream;
    WaveDev *Dp = wave_dev();
    PCX_REQUIRE(Dp != nullptr, "no current CUDA device");
    WaveDev &D = *Dp;
    std::lock_guard<std::mutex> dev_lock(D.mu);
    const int Hf = n.h * n.npart, nsteps = pcx_wave_steps(net);
    const size_t max_rows = (size_t)n.nimg * (size_t)(Hf < n.W ? Hf : n.W) * n.G + 16;
    rc = ensure_pinned(D, max_rows, n.nstep);
    if (rc < 0) return rc;
    cudaEvent_t ev[2];
    for (auto &e : ev) PCX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    PCX_CUDA(cudaMemsetAsync(n.d_prev, 0, sizeof(float) * (size_t)n.nimg * Hf * n.W, s));      // label of "step -1" is all zero
    int counts[2] = {0, 0};
    long long total = 0;
    int status = PCX_OK;
    CoderPool pool;
    pool.start(coders, n.nimg, n.nstep, true);
    for (int step = 0; step <= nsteps && status == PCX_OK; step++) {
        const int b = step & 1;
        if (step < nsteps) {
            int cnt = 0;
            status = wave_launch_step(n, step, n.d_prev, &cnt, s);
            if (status < 0) break;
            // labels of this step = DExtract2(label=true) of the symbol tensor; they are next step's DInput2 source
            int lcnt = 0;
            status = pcx_dextract_step(d_data, n.d_prev, n.nimg, n.npart, n.G, 1, n.h, n.W, step, 0, 0, n.d_order, n.h_start, &lcnt, s);
            if (status < 0) break;
            counts[b] = cnt;
            if (cnt > 0) {
                PCX_CUDA(cudaMemcpyAsync(D.pin.cdf[b], n.d_cdf, sizeof(int32_t) * (size_t)cnt * (n.nstep + 1), cudaMemcpyDeviceToHost, s));
                PCX_CUDA(cudaMemcpyAsync(D.pin.lab[b], n.d_prev, sizeof(float) * (size_t)cnt, cudaMemcpyDeviceToHost, s));
            }
            PCX_CUDA(cudaEventRecord(ev[b], s));
        }
        if (step > 0) {                             // code the previous step while the GPU runs this one
            const int pb = (step - 1) & 1;
            PCX_CUDA(cudaEventSynchronize(ev[pb]));
            const int cnt = counts[pb];
            if (cnt > 0) {
                int32_t *lab = reinterpret_cast<int32_t *>(D.pin.lab[pb]);
                for (int i = 0; i < cnt; i++) lab[i] = (int32_t)D.pin.lab[pb][i];            // symbols 0..7 stored as float
                // rows are ordered (image, cell of the window): every image has its own bitstream and its own host thread
                status = pool.run(D.pin.cdf[pb], lab, nullptr, cnt / n.nimg);
                total += cnt;
            }
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    if (n_symbols) *n_symbols = total;
    return status;
}

// One-shot encoder: all symbols are known, so every layer is evaluated over the whole tensor in one launch
// (12 x [halo, wrap, masked conv, add] instead of nsteps x 32 launches).  Each output scalar goes through the same FFMA chains
// and fold tree as in the stepwise form, the halo cells are interpolated from the same final values, and the CDF rows are
// emitted in the stepwise coding order - the bitstream is byte-identical (tests/test_gpu_codec.py).  The rows are produced in
// chunks of whole steps; the host codes chunk i (one thread per image) while chunk i+1 is computed and copied.
static std::atomic<int> g_opt_slabs{0}, g_opt_tsplit{0}, g_opt_smem{1};

int pcx_wave_set_option(const char *name, int value)
{
    if (!name) return PCX_EINVAL;
    std::atomic<int> *o = !strcmp(name, "slabs") ? &g_opt_slabs : (!strcmp(name, "tsplit") ? &g_opt_tsplit : (!strcmp(name, "smem") ? &g_opt_smem : nullptr));
    if (!o) { pcx_set_error("unknown option %s", name); return PCX_EINVAL; }
    return o->exchange(value);
}

int pcx_wave_encode_full(const pcx_wave_net *net, const float *d_data, pcx_coder *const *coders, long long *n_symbols, void *stream)
{
    int rc = wave_check(net);
    if (rc < 0) return rc;
    PCX_REQUIRE(d_data && coders, "null data / coders");
    for (int i = 0; i < net->nimg; i++) PCX_REQUIRE(coders[i] != nullptr, "null coder for image %d", i);
    const pcx_wave_net &n = *net;
    PCX_REQUIRE(n.d_lab && n.d_steptab && n.cdf_rows > 0, "one-shot encoding needs d_lab, d_steptab and cdf_rows");
    PCX_REQUIRE(n.ng == 3 && n.nstep == 8, "the one-shot encoder is built for 3 Gaussians / 8 symbols (got %d / %d)", n.ng, n.nstep);
    cudaStream_t s = (cudaStream_t)stream;
    const int Hf = n.h * n.npart, nsteps = pcx_wave_steps(net), nrep = n.nb * n.nimg;
    Bands bands;
    PCX_REQUIRE(make_bands(bands, n.wl, n.npart) == 0, "bad band description");

    // per-step cell windows and the row prefix of the coding order
    std::vector<int> tab(2 * (nsteps + 1));
    int *wfirst = tab.data(), *rowbase = tab.data() + nsteps + 1;
    int maxcnt = 0;
    rowbase[0] = 0;
    for (int st = 0; st < nsteps; st++) {
        Window w = wave_window(n.h_start, st, n.G, Hf, n.W);
        wfirst[st] = w.first;
        rowbase[st + 1] = rowbase[st] + w.count;
        if (w.count > maxcnt) maxcnt = w.count;
    }
    wfirst[nsteps] = 0;
    const int cap = n.cdf_rows / n.nimg;                      // rows per image per chunk
    PCX_REQUIRE(cap >= maxcnt, "cdf_rows %d too small for a step of %d rows x %d images", n.cdf_rows, maxcnt, n.nimg);
    WaveDev *Dp = wave_dev();
    PCX_REQUIRE(Dp != nullptr, "no current CUDA device");
    WaveDev &D = *Dp;
    std::lock_guard<std::mutex> dev_lock(D.mu);
    rc = ensure_pinned(D, (size_t)n.cdf_rows, n.nstep);
    if (rc < 0) return rc;
    // PCX_ENCODE_TRACE=1: wall-clock milestones on stderr (the stream is synchronised at each one: diagnosis only)
    static const bool enc_trace = getenv("PCX_ENCODE_TRACE") != nullptr;
    const auto trace_t0 = std::chrono::steady_clock::now();
    auto trace_mark = [&](const char *what, bool sync) {
        if (!enc_trace) return;
        if (sync) cudaStreamSynchronize(s);
        fprintf(stderr, "[pcx encode] %8.3f ms  %s\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - trace_t0).count(), what);
    };
    PCX_CUDA(cudaMemcpyAsync(n.d_steptab, tab.data(), sizeof(int) * tab.size(), cudaMemcpyHostToDevice, s));

    // ---- the network, layer by layer over the whole tensor
    const i64 in_elems = (i64)nrep * n.npart * n.G * n.layers[0].gi * (n.h + 2 * n.pad) * (n.W + 2 * n.pad);
    PCX_CUDA(cudaMemsetAsync(n.layers[0].in, 0, sizeof(float) * in_elems, s));
    {
        const i64 total = (i64)n.nimg * n.npart * n.G * n.h * n.W;
        const i64 rep_stride = (i64)n.nimg * n.npart * n.G * (n.h + 2 * n.pad) * (n.W + 2 * n.pad);
        dinput_full_kernel<<<grid_for(total, 256), 256, 0, s>>>(d_data, n.layers[0].in, bands, total, n.G, n.h, n.W, n.pad,
                                                                n.input_bias, n.nb, rep_stride);
        PCX_LAUNCHED();
    }
    // Slabs.  The symbols of wavefront step st depend only on scalars of steps <= st, so the tensor is cut
    // into consecutive step ranges [slab_lo, slab_hi): slab k runs all layers on its own scalars only, its CDF rows go out and
    // the host codes them while the device is already computing slab k + 1 - the serial host coder (~12 ns per symbol, as long
    // as all device work of a 2048x4096 image) is hidden behind the device instead of following it.  Needs the tiled kernels
    // (they take the step range); any other layer shape falls back to one slab.
    const int npl_items = Hf + n.W + n.pad - 1;                  // planes of the halo / wrap work lists (h_pstart)
    bool tiled_all = n.W % (32 * CT_CPT) == 0;
    for (int L = 0; L < n.nlayers; L++) tiled_all = tiled_all && n.layers[L].go == 3 && (n.layers[L].gi == 1 || n.layers[L].gi == 3);
    int nslab = 1;
    {
        static const char *e = getenv("PCX_WAVE_SLABS");
        // A tile of the tiled kernels is 128 cells = 128 consecutive steps wide and is recomputed by every slab it touches, so
        // slabs pay off only when a slab spans several tile widths (measured, entropy encode in ms, 1 / 2 / 4 / 8 slabs:
        // 512x1024 3.8 / 5.1 / 8.4 / 15.1; 2048x4096 34.0 / 28.8 / 30.4 / 40.2)
        const int opt = g_opt_slabs.load();
        const int want = opt > 0 ? opt : (e ? atoi(e) : (nsteps >= 1400 ? 3 : (nsteps >= 600 ? 2 : 1)));
        if (tiled_all && want > 1) nslab = want > 16 ? 16 : want;
        if (nslab > nsteps) nslab = nsteps > 0 ? nsteps : 1;
    }
    std::vector<int> slab_end(nslab);                            // first step AFTER slab k, cut at equal shares of the rows
    for (int k = 0, st = 0; k < nslab; k++) {
        const long long target = (long long)rowbase[nsteps] * (k + 1) / nslab;
        while (st < nsteps && (rowbase[st] < target || st <= (k ? slab_end[k - 1] : 0))) st++;
        slab_end[k] = k == nslab - 1 ? nsteps : st;
    }
    if (!D.s_copy) {                                             // CDF rows leave on their own stream, next to the next slab
        PCX_CUDA(cudaStreamCreateWithFlags(&D.s_copy, cudaStreamNonBlocking));
        for (auto &e : D.slab_ev) PCX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t s_copy = D.s_copy;
    cudaEvent_t *slab_ev = D.slab_ev;
    for (int L = 0; L < n.nlayers; L++) {
        const pcx_wave_layer &l = n.layers[L];
        const i64 out_elems = (i64)nrep * n.npart * n.G * l.go * (n.h + 2 * l.pad_out) * (n.W + 2 * l.pad_out);
        PCX_CUDA(cudaMemsetAsync(l.out, 0, sizeof(float) * out_elems, s));
    }
    for (int slab = 0; slab < nslab; slab++) {
    const int slab_lo = nslab == 1 ? 0 : (slab ? slab_end[slab - 1] : 0), slab_hi = nslab == 1 ? 0x7fffffff : slab_end[slab];
    // halo / wrap cells this slab can read: planes [slab_lo - G - 4, slab_hi) (earlier ones are final, later ones not needed yet)
    int pl_a = nslab == 1 ? 0 : slab_lo - n.G - 4, pl_b = nslab == 1 ? npl_items : slab_hi;
    pl_a = pl_a < 0 ? 0 : (pl_a > npl_items ? npl_items : pl_a);
    pl_b = pl_b > npl_items ? npl_items : pl_b;
    const int item0 = n.h_pstart[pl_a], nitems = n.h_pstart[pl_b] - n.h_pstart[pl_a];
    for (int L = 0; L < n.nlayers; L++) {
        const pcx_wave_layer &l = n.layers[L];
        const int Ci = n.G * l.gi, Co = n.G * l.go;
        if (nitems > 0) {
            const i64 total = (i64)nitems * Ci * nrep;
            for (int kind = 0; kind < 2; kind++) {
                ctx_pad_kernel<<<grid_for(total, 256), 256, 0, s>>>(l.in, bands, n.d_band, n.d_row, n.d_col, n.d_tw,
                                                                    (const int4 *)n.d_items, item0, nitems, nrep, l.gi, Ci, n.h, n.W,
                                                                    n.pad, 0, 1, kind);
                PCX_LAUNCHED();
            }
        }
        static const bool no_smem_form = getenv("PCX_CTX_NO_SMEM") != nullptr;
        const size_t cs_bytes = (size_t)n.G * l.gi * 25 * 2 * sizeof(float4) + (size_t)n.G * l.gi * (CS_ROWS + 4) * CS_COLS * sizeof(float);
        // one block per SM (211 KB of shared memory): worth it from two full waves of blocks on (2048x4096: 768 blocks, 40.7 ->
        // 36.5 ms for the whole entropy encode; a single 512x1024 image is 48 blocks and stays on the L1-resident form)
        const i64 cs_blocks = (i64)(n.W / (32 * CT_CPT)) * ceil_div(n.h, CS_ROWS) * n.npart * nrep;
        // how many blocks share the channel-group pairs of a tile: fewest waves x (staging + pairs per block), measured ~10 us to
        // stage a window and ~27 us per pair (2048x4096: 1.00 ms per layer with 768 blocks x 7 pairs)
        int tsplit = 1;
        {
            const int npair = (n.G + 1) / 2, sms = pcx_sm_count();
            double best = 1e30;
            for (int t = 1; t <= npair; t++) {
                const double cost = (double)ceil_div(cs_blocks * t, sms) * (10.0 + 27.0 * ceil_div(npair, t));
                if (cost < best - 1e-9) { best = cost; tsplit = t; }
            }
            static const char *e = getenv("PCX_CTX_TSPLIT");
            if (e && atoi(e) >= 1 && atoi(e) <= npair) tsplit = atoi(e);
            const int opt = g_opt_tsplit.load();
            if (opt >= 1 && opt <= npair) tsplit = opt;
        }
        if (!no_smem_form && g_opt_smem.load() != 0 && l.go == 3 && n.W % (32 * CT_CPT) == 0 && (l.gi == 1 || l.gi == 3) && n.pad == 2 && cs_bytes <= 220 * 1024 &&
            (i64)nrep * tsplit <= 65535) {
            // shared-memory form: the block stages its 5x5 input window once and loops over the channel groups
            // measured at 2048x4096 (ms per layer): x1 1.00, x2 1.12, runtime tree 1.25-1.42, full unroll 2.53
            static const int unroll = [] {
                const char *e = getenv("PCX_CTX_UNROLL");
                const int u = e ? atoi(e) : 1;
                return (u == 2 || u == 8 || u == 9) ? u : 1;
            }();
            static PcxDeviceOnce cs_once;
            PCX_ONCE_PER_DEVICE(cs_once) {
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<1, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_smem_kernel<3, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
            }
            const int ntile = n.W / (32 * CT_CPT);
            dim3 grid((unsigned)(ntile * ceil_div(n.h, CS_ROWS)), (unsigned)n.npart, (unsigned)(nrep * tsplit));
#define PCX_CS_LAUNCH(GI_, U_)                                                                                                              \
    ctx_conv_smem_kernel<GI_, U_><<<grid, 64 * CS_ROWS, cs_bytes, s>>>(l.in, l.weight, l.bias, l.act, l.add, l.out, n.nimg, n.npart, n.G, n.h, \
                                                                       n.W, n.pad, l.pad_out, l.constrain, bands, slab_lo, slab_hi, tsplit)
            if (l.gi == 1) {
                if (unroll == 1) PCX_CS_LAUNCH(1, 1); else if (unroll == 2) PCX_CS_LAUNCH(1, 2); else if (unroll == 8) PCX_CS_LAUNCH(1, 8); else PCX_CS_LAUNCH(1, 9);
            } else {
                if (unroll == 1) PCX_CS_LAUNCH(3, 1); else if (unroll == 2) PCX_CS_LAUNCH(3, 2); else if (unroll == 8) PCX_CS_LAUNCH(3, 8); else PCX_CS_LAUNCH(3, 9);
            }
#undef PCX_CS_LAUNCH
            PCX_LAUNCHED();
            continue;
        }
        if (l.go == 3 && n.W % (32 * CT_CPT) == 0 && (l.gi == 1 || l.gi == 3)) {
            // tiled throughput form, residual add fused into the store
            const int ntile = n.W / (32 * CT_CPT);
            dim3 grid((unsigned)ceil_div((i64)Hf * ntile, CT_WARPS), (unsigned)n.G, (unsigned)nrep);
            const size_t smem = (size_t)Ci * 25 * sizeof(float4);
            if (smem > 48 * 1024) {                       // 48 channel groups (valid_dim 192): 57.6 KB of weight rows
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<3, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                PCX_CUDA(cudaFuncSetAttribute(ctx_conv_tiled_kernel<3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            }
            static const int ct_unroll = getenv("PCX_CT_UNROLL") ? atoi(getenv("PCX_CT_UNROLL")) : 0;
#define PCX_CT_LAUNCH(GI_, U_)                                                                                                              \
    ctx_conv_tiled_kernel<GI_, U_><<<grid, 32 * CT_WARPS, smem, s>>>(l.in, l.weight, l.bias, l.act, l.add, l.out, n.nimg, n.npart, n.G, n.h,  \
                                                                     n.W, n.pad, l.pad_out, l.constrain, bands, slab_lo, slab_hi)
            if (l.gi == 1) {
                if (ct_unroll == 2) PCX_CT_LAUNCH(1, 2); else if (ct_unroll == 4) PCX_CT_LAUNCH(1, 4); else PCX_CT_LAUNCH(1, 0);
            } else {
                if (ct_unroll == 2) PCX_CT_LAUNCH(3, 2); else if (ct_unroll == 4) PCX_CT_LAUNCH(3, 4); else PCX_CT_LAUNCH(3, 0);
            }
#undef PCX_CT_LAUNCH
            PCX_LAUNCHED();
            continue;
        }
        const i64 nscalars = (i64)nrep * l.go * n.G * Hf * n.W;
        PCX_REQUIRE(nscalars * 32 / 256 < 0x7fffffffll, "tensor too large for one launch");
        const int blocks = grid_for(nscalars * 32, 256);
        if (l.gi == 1)
            ctx_conv_kernel<1, true><<<blocks, 256, 0, s>>>(l.in, l.weight, l.bias, l.act, l.out, nullptr, 0, 0, n.nimg, nscalars, n.npart,
                                                            n.G, l.go, n.h, n.W, n.pad, l.pad_out, l.constrain, 0, bands);
        else if (l.gi == 3)
            ctx_conv_kernel<3, true><<<blocks, 256, 0, s>>>(l.in, l.weight, l.bias, l.act, l.out, nullptr, 0, 0, n.nimg, nscalars, n.npart,
                                                            n.G, l.go, n.h, n.W, n.pad, l.pad_out, l.constrain, 0, bands);
        else
            PCX_REQUIRE(false, "input channels per group must be 1 or 3 (got %d)", l.gi);
        PCX_LAUNCHED();
        if (l.add) {
            const i64 total = (i64)nrep * n.npart * Co * n.h * n.W;
            ctx_add_full_kernel<<<grid_for(total, 256), 256, 0, s>>>(l.out, l.add, bands, total, Co, n.h, n.W, l.pad_out);
            PCX_LAUNCHED();
        }
    }
    PCX_CUDA(cudaEventRecord(slab_ev[slab], s));
    trace_mark("slab computed", true);
    }

    // ---- CDF rows in coding order, chunk by chunk, host coding pipelined behind the device
    const pcx_wave_layer &last = n.layers[n.nlayers - 1];
    cudaEvent_t ev[2];
    for (auto &e : ev) PCX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CoderPool pool;
    pool.start(coders, n.nimg, n.nstep, true);
    trace_mark("coder threads started", false);
    int status = PCX_OK, per_chunk[2] = {0, 0}, s0 = 0, chunk = 0;
    long long total_rows = 0;
    auto code_chunk = [&](int b) -> int {
        cudaError_t e = cudaEventSynchronize(ev[b]);
        if (e != cudaSuccess) { pcx_set_error("cudaEventSynchronize -> %s", cudaGetErrorString(e)); return PCX_ECUDA; }
        trace_mark("chunk on the host", false);
        if (per_chunk[b] <= 0) return PCX_OK;
        total_rows += (long long)per_chunk[b] * n.nimg;
        const int r = pool.run(D.pin.cdf[b], reinterpret_cast<int32_t *>(D.pin.lab[b]), nullptr, per_chunk[b]);
        trace_mark("chunk coded", false);
        return r;
    };
    int cur_slab = 0;
    while (s0 < nsteps && status == PCX_OK) {
        while (s0 >= slab_end[cur_slab]) cur_slab++;
        const int lim = slab_end[cur_slab];                    // a chunk never crosses into a slab that is still being computed
        int s1 = s0 + 1;
        while (s1 < lim && rowbase[s1 + 1] - rowbase[s0] <= cap) s1++;
        const int per = rowbase[s1] - rowbase[s0], b = chunk & 1;
        per_chunk[b] = per;
        PCX_CUDA(cudaStreamWaitEvent(s_copy, slab_ev[cur_slab], 0));
        if (per > 0) {
            const i64 rows = (i64)per * n.nimg;
            gmm_ordered_kernel<<<grid_for(rows, 128), 128, 0, s_copy>>>(last.out, d_data, n.d_order, n.d_steptab, n.d_steptab + nsteps + 1,
                                                                        s0, s1, n.nimg, n.npart, n.G, last.go, n.h, n.W, n.ng, n.nstep,
                                                                        n.gmm_bias, n.gmm_total, n.gmm_beta, n.d_cdf, n.d_lab);
            PCX_LAUNCHED();
            PCX_CUDA(cudaMemcpyAsync(D.pin.cdf[b], n.d_cdf, sizeof(int32_t) * (size_t)rows * (n.nstep + 1), cudaMemcpyDeviceToHost, s_copy));
            PCX_CUDA(cudaMemcpyAsync(D.pin.lab[b], n.d_lab, sizeof(int32_t) * (size_t)rows, cudaMemcpyDeviceToHost, s_copy));
        }
        PCX_CUDA(cudaEventRecord(ev[b], s_copy));
        if (chunk >= 1) status = code_chunk(b ^ 1);           // code chunk c-1 while chunk c is computed and copied
        s0 = s1;
        chunk++;
    }
    if (status == PCX_OK && chunk >= 1) status = code_chunk((chunk - 1) & 1);
    // the caller's stream observes the copy stream's work (scratch buffers may be reused right after this call)
    if (chunk >= 1) {
        PCX_CUDA(cudaEventRecord(ev[0], s_copy));
        PCX_CUDA(cudaStreamWaitEvent(s, ev[0], 0));
        PCX_CUDA(cudaStreamSynchronize(s_copy));
    }
    for (auto &e : ev) cudaEventDestroy(e);
    if (n_symbols) *n_symbols = total_rows;
    return status;
}

static std::atomic<int> g_wave_fused{2};

int pcx_wave_set_fused(int on)
{
    return g_wave_fused.exchange(on < 0 ? 0 : (on > 2 ? 2 : on));
}

// Decoder loop on the fused step kernel: per step ONE cooperative launch, a stream synchronisation, the host range decoder
// (one thread per image); CDF rows and decoded symbols cross PCIe through mapped pinned memory, no copy calls.
static int wave_decode_fused(const pcx_wave_net &n, pcx_coder *const *coders, long long *n_symbols, cudaStream_t s)
{
    const int Hf = n.h * n.npart, nsteps = pcx_wave_steps(&n), nrep = n.nb * n.nimg;
    const size_t max_rows = (size_t)n.nimg * (size_t)(Hf < n.W ? Hf : n.W) * n.G + 16;
    WaveDev *Dp = wave_dev();
    PCX_REQUIRE(Dp != nullptr, "no current CUDA device");
    WaveDev &D = *Dp;
    std::lock_guard<std::mutex> dev_lock(D.mu);
    int rc = ensure_pinned(D, max_rows, n.nstep);
    if (rc < 0) return rc;
    if (D.bar == nullptr) PCX_CUDA(cudaMalloc((void **)&D.bar, 64));
    PCX_CUDA(cudaMemsetAsync(D.bar, 0, 64, s));
    StepNet d;
    d.nlayers = n.nlayers; d.nb = n.nb; d.nimg = n.nimg; d.npart = n.npart; d.G = n.G; d.h = n.h; d.W = n.W; d.pad = n.pad;
    d.nstep = n.nstep; d.ng = n.ng;
    d.gmm_bias = n.gmm_bias; d.gmm_total = n.gmm_total; d.gmm_beta = n.gmm_beta; d.input_bias = n.input_bias;
    PCX_REQUIRE(make_bands(d.bands, n.wl, n.npart) == 0, "bad band description");
    d.hband = n.d_band; d.hrow = n.d_row; d.hcol = n.d_col; d.htw = n.d_tw; d.order = n.d_order;
    float *dev_prev = nullptr;
    int *dev_cdf = nullptr;
    PCX_CUDA(cudaHostGetDevicePointer((void **)&dev_prev, D.pin.lab[0], 0));
    PCX_CUDA(cudaHostGetDevicePointer((void **)&dev_cdf, D.pin.cdf[0], 0));
    d.prev = dev_prev; d.cdf = dev_cdf; d.bar = D.bar;
    // PCX_WAVE_TRACE=<file>: per-step device timestamps (block 0) + host wall clock of the loop, written as text
    const char *trace_path = getenv("PCX_WAVE_TRACE");
    unsigned long long *h_dbg = nullptr;
    struct HostGuard {                           // the trace buffer is released on every exit path
        unsigned long long **p;
        ~HostGuard() { if (*p) cudaFreeHost(*p); }
    } dbg_guard{&h_dbg};
    std::vector<double> host_us;
    d.dbg = nullptr;
    if (trace_path) {
        PCX_CUDA(cudaHostAlloc((void **)&h_dbg, sizeof(unsigned long long) * 32 * (size_t)nsteps, cudaHostAllocMapped));
        memset(h_dbg, 0, sizeof(unsigned long long) * 32 * (size_t)nsteps);
        PCX_CUDA(cudaHostGetDevicePointer((void **)&d.dbg, h_dbg, 0));
        host_us.resize(3 * (size_t)nsteps);
    }
    // channels-last scratch: input of layer 0 (cp = 8*ceil(G/8)*gi0) and one output per layer (cp = 8*ceil(G/8)*3)
    const int G8 = (n.G + 7) / 8 * 8;
    const i64 planes = (i64)nrep * n.npart;
    const i64 in_elems = planes * (n.h + 2 * n.pad) * (n.W + 2 * n.pad);
    std::vector<i64> off(n.nlayers + 2);
    off[0] = 0;
    off[1] = in_elems * G8 * n.layers[0].gi;
    for (int L = 0; L < n.nlayers; L++)
        off[L + 2] = off[L + 1] + planes * (n.h + 2 * n.layers[L].pad_out) * (n.W + 2 * n.layers[L].pad_out) * G8 * 3;
    const int ncell_total = n.h_start[Hf + n.W - 1];                              // valid cells of one image
    const size_t cell_off = (sizeof(float) * (size_t)off[n.nlayers + 1] + 255) / 256 * 256;
    const size_t need_bytes = cell_off + sizeof(int4) * (size_t)ncell_total + 256;
    if (need_bytes > D.step_scratch_bytes) {
        if (D.step_scratch) cudaFree(D.step_scratch);
        D.step_scratch = nullptr;
        D.step_scratch_bytes = 0;
        PCX_CUDA(cudaMalloc((void **)&D.step_scratch, need_bytes));
        D.step_scratch_bytes = need_bytes;
    }
    PCX_CUDA(cudaMemsetAsync(D.step_scratch, 0, need_bytes, s));
    for (int L = 0; L < n.nlayers; L++) {
        const pcx_wave_layer &l = n.layers[L];
        StepLayer &sl = d.L[L];
        sl.weight = l.weight; sl.bias = l.bias; sl.act = l.act;
        sl.in = D.step_scratch + off[L];
        sl.out = D.step_scratch + off[L + 1];
        sl.add = nullptr;
        if (l.add) {
            for (int M = 0; M < L; M++)
                if (n.layers[M].out == l.add) sl.add = D.step_scratch + off[M + 1];
            PCX_REQUIRE(sl.add != nullptr, "layer %d: the residual source must be the output of an earlier layer", L);
            PCX_REQUIRE(n.layers[L].pad_out == n.pad, "layer %d: residual add on an unpadded output", L);
        }
        sl.gi = l.gi; sl.pad_out = l.pad_out; sl.constrain = l.constrain;
        sl.cp_in = G8 * l.gi; sl.cp_out = G8 * 3;
        PCX_REQUIRE(L == 0 || (l.in == n.layers[L - 1].out && l.gi == 3 && n.layers[L - 1].pad_out == n.pad),
                    "layer %d must read the padded output of layer %d", L, L - 1);
    }
    d.sym_nchw = n.layers[0].in;
    int4 *d_cell = reinterpret_cast<int4 *>(reinterpret_cast<char *>(D.step_scratch) + cell_off);
    step_cellinfo_kernel<<<ceil_div(ncell_total, 256), 256, 0, s>>>(n.d_order, d_cell, ncell_total, n.h, n.W);
    PCX_LAUNCHED();
    d.cell = d_cell;
    {
        float f = (float)(1 - 1e-6);
        if ((double)f < 1 - 1e-6) f = nextafterf(f, 2.0f);
        d.halo_one = f;
    }
    PCX_CUDA(cudaMemsetAsync(n.layers[0].in, 0, sizeof(float) * (size_t)nrep * n.npart * n.G * (n.h + 2 * n.pad) * (n.W + 2 * n.pad), s));

    const int threads = STEP_THREADS;
    const size_t smem = sizeof(float) * 2 * (4 * (size_t)n.G * 3 * 25 + 8);
    PCX_CUDA(cudaFuncSetAttribute(wave_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PCX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wave_step_kernel, threads, smem));
    PCX_REQUIRE(per_sm >= 1, "wave_step_kernel does not fit on an SM");
    const int max_grid = pcx_sm_count() * per_sm;
    const int nplanes = Hf + n.W - 1;
    if ((size_t)(Hf + n.W) > D.d_start_cap) {               // cached: no allocation (and no leak on an early return) per call
        if (D.d_start) cudaFree(D.d_start);
        D.d_start = nullptr;
        D.d_start_cap = 0;
        PCX_CUDA(cudaMalloc((void **)&D.d_start, sizeof(int) * (size_t)(Hf + n.W)));
        D.d_start_cap = (size_t)(Hf + n.W);
    }
    int *d_start = D.d_start;
    PCX_CUDA(cudaMemcpyAsync(d_start, n.h_start, sizeof(int) * (size_t)(Hf + n.W), cudaMemcpyHostToDevice, s));

    CoderPool pool;
    pool.start(coders, n.nimg, n.nstep, false);
    long long total = 0;
    unsigned bar_base = 0;
    FILE *dump = getenv("PCX_WAVE_DUMP") ? fopen(getenv("PCX_WAVE_DUMP"), "w") : nullptr;
    int pfirst = 0, pcount = 0;
    for (int step = 0; step < nsteps; step++) {
        int p0 = step - n.G + 1 < 0 ? 0 : step - n.G + 1;
        int p1 = step < nplanes - 1 ? step + 1 : nplanes;                  // planes [p0, p1) (entropy_conv_cuda_v2.cu:389-391)
        if (p0 > p1) p0 = p1;
        int np = p1 - p0;
        int first = n.h_start[p0], count = n.h_start[p1] - n.h_start[p0];
        // runs of at most S cells per (net image, plane); block i takes runs i, i + grid, ...  S (a multiple of the warps per
        // block) minimises rounds-per-block x cells-per-warp, the critical path of a layer
        const int wpb = threads / 32;
        int maxcells = 1;
        for (int q = p0; q < p1; q++) maxcells = n.h_start[q + 1] - n.h_start[q] > maxcells ? n.h_start[q + 1] - n.h_start[q] : maxcells;
        maxcells *= n.nimg;
        int S = wpb, nchunk = 0;
        long best = -1;
        for (int cand = wpb; cand < maxcells + wpb; cand += wpb) {
            int chunks = 0;
            for (int q = p0; q < p1; q++) chunks += n.nb * ceil_div((i64)(n.h_start[q + 1] - n.h_start[q]) * n.nimg, cand);
            const long cost = (long)ceil_div(chunks, max_grid) * (cand / wpb);
            if (best < 0 || cost <= best) { best = cost; S = cand; nchunk = chunks; }      // ties: the longer run (fewer weight stagings)
        }
        int grid = nchunk < max_grid ? nchunk : max_grid;
        if (grid < 1) grid = 1;
        const int *dstart = d_start;
        void *args[] = {(void *)&d, (void *)&dstart, (void *)&step, (void *)&p0, (void *)&np, (void *)&S, (void *)&nchunk, (void *)&pfirst,
                        (void *)&pcount, (void *)&bar_base};
        auto t0 = std::chrono::steady_clock::now();
        PCX_CUDA(cudaLaunchCooperativeKernel((const void *)wave_step_kernel, dim3(grid), dim3(threads), args, smem, s));
        PCX_LAUNCHED();
        bar_base += (unsigned)(1 + n.nlayers) * (unsigned)grid;
        auto t1 = std::chrono::steady_clock::now();
        PCX_CUDA(cudaStreamSynchronize(s));
        auto t2 = std::chrono::steady_clock::now();
        if (count > 0) {
            rc = pool.run(D.pin.cdf[0], nullptr, D.pin.lab[0], count);
            if (dump) {                                 // PCX_WAVE_DUMP=<file>: same text as the dataflow engine writes (debugging)
                for (int im = 0; im < n.nimg; im++)
                    for (int r = 0; r < count; r++) {
                        const int32_t *row = D.pin.cdf[0] + ((size_t)im * count + r) * 9;
                        fprintf(dump, "%d %d %d |", im, step, r);
                        for (int j = 1; j < 8; j++) fprintf(dump, " %d", row[j]);
                        fprintf(dump, " | %d\n", (int)D.pin.lab[0][(size_t)im * count + r]);
                    }
            }
            if (rc < 0) { if (dump) fclose(dump); return rc; }
            total += (long long)count * n.nimg;
        }
        if (h_dbg) {
            auto t3 = std::chrono::steady_clock::now();
            host_us[3 * (size_t)step + 0] = std::chrono::duration<double, std::micro>(t1 - t0).count();
            host_us[3 * (size_t)step + 1] = std::chrono::duration<double, std::micro>(t2 - t1).count();
            host_us[3 * (size_t)step + 2] = std::chrono::duration<double, std::micro>(t3 - t2).count();
        }
        pfirst = first;
        pcount = count;
    }
    if (dump) fclose(dump);
    // the last step's symbols never pass through the network: one more DInput2 (pseudo_codec.py:159)
    rc = pcx_dinput_step(dev_prev, n.layers[0].in, n.nimg, n.npart, n.G, n.h, n.W, n.pad, n.input_bias, n.nb, nsteps, n.d_order, n.h_start, s);
    if (rc < 0) return rc;
    PCX_CUDA(cudaStreamSynchronize(s));
    if (h_dbg) {
        FILE *f = fopen(trace_path, "w");
        if (f) {
            fprintf(f, "# step count launch_us sync_us code_us | device ns: dinput, layer 0..%d, gmm (block 0, globaltimer)\n", n.nlayers - 1);
            for (int st = 0; st < nsteps; st++) {
                Window w = wave_window(n.h_start, st, n.G, Hf, n.W);
                fprintf(f, "%d %d %.1f %.1f %.1f |", st, w.count, host_us[3 * (size_t)st], host_us[3 * (size_t)st + 1], host_us[3 * (size_t)st + 2]);
                for (int i = 1; i <= 2 + n.nlayers; i++)
                    fprintf(f, " %lld", (long long)(h_dbg[(size_t)st * 32 + i] - h_dbg[(size_t)st * 32 + i - 1]));
                fprintf(f, "\n");
            }
            fclose(f);
        }
        cudaFreeHost(h_dbg);
        h_dbg = nullptr;
    }
    if (n_symbols) *n_symbols = total;
    return PCX_OK;
}

int pcx_wave_decode(const pcx_wave_net *net, pcx_coder *const *coders, long long *n_symbols, void *stream)
{
    int rc = wave_check(net);
    if (rc < 0) return rc;
    PCX_REQUIRE(coders, "null coders");
    for (int i = 0; i < net->nimg; i++) PCX_REQUIRE(coders[i] != nullptr, "null coder for image %d", i);
    const pcx_wave_net &n = *net;
    cudaStream_t s = (cudaStream_t)stream;
    if (g_wave_fused.load()) {
        bool ok = n.nb == 3 && n.ng == 3 && n.nstep == 8;
        for (int L = 0; L < n.nlayers; L++) ok = ok && n.layers[L].go == 3 && (n.layers[L].gi == 1 || n.layers[L].gi == 3);
        if (ok && g_wave_fused.load() == 2) {
            bool unsupported = false;
            rc = pcx_wave_decode_flow(n, coders, n_symbols, s, &unsupported);
            if (!unsupported) return rc;
        }
        if (ok) return wave_decode_fused(n, coders, n_symbols, s);
    }
    const int Hf = n.h * n.npart, nsteps = pcx_wave_steps(net);
    const size_t max_rows = (size_t)n.nimg * (size_t)(Hf < n.W ? Hf : n.W) * n.G + 16;
    WaveDev *Dp = wave_dev();
    PCX_REQUIRE(Dp != nullptr, "no current CUDA device");
    WaveDev &D = *Dp;
    std::lock_guard<std::mutex> dev_lock(D.mu);
    rc = ensure_pinned(D, max_rows, n.nstep);
    if (rc < 0) return rc;
    PCX_CUDA(cudaMemsetAsync(n.d_prev, 0, sizeof(float) * (size_t)n.nimg * Hf * n.W, s));
    long long total = 0;
    CoderPool pool;
    pool.start(coders, n.nimg, n.nstep, false);
    for (int step = 0; step < nsteps; step++) {
        int cnt = 0;
        rc = wave_launch_step(n, step, n.d_prev, &cnt, s);
        if (rc < 0) return rc;
        if (cnt > 0) {
            PCX_CUDA(cudaMemcpyAsync(D.pin.cdf[0], n.d_cdf, sizeof(int32_t) * (size_t)cnt * (n.nstep + 1), cudaMemcpyDeviceToHost, s));
            PCX_CUDA(cudaStreamSynchronize(s));
            rc = pool.run(D.pin.cdf[0], nullptr, D.pin.lab[0], cnt / n.nimg);
            if (rc < 0) return rc;
            PCX_CUDA(cudaMemcpyAsync(n.d_prev, D.pin.lab[0], sizeof(float) * (size_t)cnt, cudaMemcpyHostToDevice, s));
            total += cnt;
        }
    }
    // the last step's symbols never pass through the network: one more DInput2 puts them into the padded input buffer,
    // which then holds every decoded symbol (+ input_bias) - pseudo_codec.py:159
    rc = pcx_dinput_step(n.d_prev, n.layers[0].in, n.nimg, n.npart, n.G, n.h, n.W, n.pad, n.input_bias, n.nb, nsteps, n.d_order, n.h_start, s);
    if (rc < 0) return rc;
    PCX_CUDA(cudaStreamSynchronize(s));
    if (n_symbols) *n_symbols = total;
    return PCX_OK;
}

}  // extern "C"
