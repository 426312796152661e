// pcx_tile_nhwc.cu - the fused tile gathers that feed / drain the tensor-core convolutions.
//
//   pcx_slice_pad_nhwc : ERP image (N,C,H,W) NCHW  ->  halo-padded band tiles [N*npart][h+2p][pitch][C] (channels last)
//                        = SphereSlice + PseudoPadV2 + layout change in ONE pass.  Every tile cell is computed
//                        straight from the ERP row it depends on: interior cells are the 4-tap Catmull-Rom
//                        resample (sphere_slice_cuda.cu:87-116); halo cells are the 2-tap interpolation
//                        (pseudo_pad.cu:57-79) of two resampled values of the neighbour band, evaluated on the fly
//                        with the same expression shapes, so the result is bit-identical to running the three
//                        reference kernels one after the other.
//   pcx_uslice_nhwc    : band tiles (channels last, any pitch / origin) -> ERP image NCHW
//                        = SphereUslice (sphere_uslice_cuda.cu:73-99) + layout change.
//
// Both are HBM-bound gathers.  A CTA owns (plane, tile row, 32 channels, a chunk of columns): it stages the
// circular span of source columns it needs in shared memory with coalesced loads along the source's
// contiguous axis, then every warp produces outputs with the lanes along the DESTINATION's contiguous axis
// (channels for the tiles, longitude for the ERP image), so both sides of the transpose move full 128-byte lines.
// Shared-memory pitches are odd so the transposed reads are bank-conflict free.
#include "pcx_common.cuh"
#include <stdlib.h>

namespace {

constexpr int CB = 32;            // channels per CTA (= one 128-byte channels-last segment)
constexpr int NTHREADS = 256;
constexpr int NWARPS = NTHREADS / 32;

__device__ __forceinline__ int wrap_mod(int v, int n)      // v in (-n, 2n)
{
    if (v < 0) v += n;
    else if (v >= n) v -= n;
    return v;
}
__device__ __forceinline__ int circ_dist(int from, int to, int n)   // from, to in [0, n)
{
    int d = to - from;
    return d < 0 ? d + n : d;
}

// 4-byte asynchronous global -> shared copy (LDGSTS): no register round trip, the warp keeps issuing
__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16(float *dst_smem, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}

__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ slice + pad
struct SlicePadParams {
    int N, C, h, W, pad, out_pitch;
    int xb[PCX_MAX_PART];   // logical tile columns per CTA, per band (narrow polar bands span more ERP columns per column)
    int xoff[PCX_MAX_PART + 1];   // prefix sum of the bands' column-chunk counts (persistent kernel's super-tile index)
    int xtotal;
    int xchunks;       // column chunks per row (max over bands)
    int cchunks;
    int scap;          // shared-memory row capacity in floats (odd pitch = scap | 1)
    int zero_invalid;  // write zeros to the columns >= wl + 2 pad
};

constexpr int XB_MAX = 128;

// Per-column gather recipe, built once per CTA and broadcast from shared memory in the inner loop.
struct ColRecipe {
    float4 wa;         // Catmull-Rom weights of the (first) resampled value
    float4 wb;         // ... of the second one (halo rows only)
    int oa, ob;        // offsets of tap 0 inside the staged span
    float t;           // halo interpolation weight
    int pad_;
};

__global__ void __launch_bounds__(NTHREADS) slice_pad_nhwc_kernel(const float *__restrict__ erp, float *__restrict__ out,
                                                                  const int *__restrict__ stab, const float4 *__restrict__ swt,
                                                                  const int *__restrict__ hband, const int *__restrict__ hrow,
                                                                  const int *__restrict__ hcol, const float *__restrict__ htw,
                                                                  Bands bands, SlicePadParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    ColRecipe *rec = reinterpret_cast<ColRecipe *>(smem_raw);                       // [XB_MAX]
    float *sm = reinterpret_cast<float *>(smem_raw + XB_MAX * sizeof(ColRecipe));   // [CB][spitch]
    const int spitch = (P.scap + 3) | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, pad = P.pad, C = P.C;
    const int OH = h + 2 * pad;

    // grid = (column chunks, channel chunks * tile rows, planes)
    const int xc = blockIdx.x;
    const int cc = blockIdx.y % (unsigned)P.cchunks;
    const int y = blockIdx.y / (unsigned)P.cchunks;
    const i64 plane = blockIdx.z;
    const int g = blockIdx.z % (unsigned)npart;
    const i64 n = blockIdx.z / (unsigned)npart;
    const int wl = bands.wl[g];
    const int c0 = cc * CB;
    const int nc = min(CB, C - c0);
    const int xb = P.xb[g];
    const int x0 = xc * xb;
    const int x1 = min(x0 + xb, wl);              // logical columns [x0, x1) of the band

    float *orow = out + ((plane * OH + y) * (i64)P.out_pitch) * C + c0;

    // ---- columns beyond the band: zeros only (PseudoPad leaves them 0; pseudo_pad.cu:39-54)
    if (P.zero_invalid) {
        // the invalid physical columns [wl + 2 pad, pitch) are split evenly over the row's CTAs
        const int zb = (P.out_pitch - wl - 2 * pad + P.xchunks - 1) / P.xchunks;
        const int X0 = wl + 2 * pad + xc * zb, X1 = min(X0 + zb, P.out_pitch);
        for (int X = X0 + warp; X < X1; X += NWARPS)
            if (lane < nc) orow[(i64)X * C + lane] = 0.f;
    }
    if (x0 >= wl) return;

    // ---- which ERP row feeds this tile row, and through which band's resampling table
    int sb, erow, hr = -1;
    if (y >= pad && y < pad + h) {
        sb = g;
        erow = g * h + (y - pad);
    } else {
        const int s = y < pad ? 0 : 1, r = y < pad ? y : y - pad - h;
        hr = (g * 2 + s) * pad + r;
        sb = hband[hr];
        erow = sb * h + hrow[hr];
    }
    const int wsrc = bands.wl[sb];
    const int *tab = stab + (i64)sb * W;
    const float4 *wtab = swt + (i64)sb * W;

    // ---- circular span of ERP columns needed by logical columns [x0, x1); small images stage the whole row
    int base = 0, span = W;
    if (W > P.scap) {
        int first, last;
        if (hr < 0) {
            first = tab[x0];
            last = tab[x1 - 1];
        } else {
            first = tab[hcol[(i64)hr * W + x0]];
            int q = hcol[(i64)hr * W + x1 - 1];
            int q1 = (q + 1 == wsrc) ? 0 : q + 1;
            last = tab[q1];
        }
        base = wrap_mod(first - 1, W);
        span = circ_dist(base, wrap_mod(last + 2, W), W) + 1;
        if (span > P.scap) __trap();              // host sizing guarantees this cannot happen
    } else {
        span = W + 3;                             // + the three wrapped columns a tap at W-1 reaches
    }

    // ---- per-column recipes
    if (threadIdx.x < x1 - x0) {
        const int x = x0 + threadIdx.x;
        ColRecipe r;
        if (hr < 0) {
            r.wa = wtab[x];
            int o = tab[x] - 1 - base;
            r.oa = o < 0 ? o + W : o;
            r.wb = r.wa; r.ob = r.oa; r.t = 0.f;
        } else {
            const i64 e = (i64)hr * W + x;
            const int q = hcol[e];
            const int q1 = (q + 1 == wsrc) ? 0 : q + 1;
            r.wa = wtab[q];
            r.wb = wtab[q1];
            int o = tab[q] - 1 - base;
            r.oa = o < 0 ? o + W : o;
            o = tab[q1] - 1 - base;
            r.ob = o < 0 ? o + W : o;
            r.t = htw[e];
        }
        r.pad_ = 0;
        rec[threadIdx.x] = r;
    }

    // ---- stage: sm[c][i] = erp[n][c0+c][erow][(base + i) mod W], lanes along i
    const float *src = erp + ((n * C + c0) * (i64)(h * npart) + erow) * W;
    const i64 cstride = (i64)h * npart * W;
    const int n1 = min(span, W - base);           // columns before the longitude wrap
    for (int c = warp; c < nc; c += NWARPS) {
        const float *sr = src + c * cstride + base;
        float *dr = sm + c * spitch;
        for (int i = lane; i < n1; i += 32) cp_async4(dr + i, sr + i);
        for (int i = n1 + lane; i < span; i += 32) cp_async4(dr + i, sr + i - W);
    }
    cp_async_wait_all();
    __syncthreads();
    if (lane >= nc) return;

    const float *mine = sm + lane * spitch;
    float *op = orow + (i64)(pad + x0 + warp) * C + lane;
    const int nx = x1 - x0;
    if (hr < 0) {
        for (int xi = warp; xi < nx; xi += NWARPS, op += (i64)NWARPS * C) {
            const int o = rec[xi].oa;
            const float4 w = rec[xi].wa;
            *op = tap4_ref<true>(w, mine[o], mine[o + 1], mine[o + 2], mine[o + 3]);
        }
    } else {
        for (int xi = warp; xi < nx; xi += NWARPS, op += (i64)NWARPS * C) {
            const ColRecipe r = rec[xi];
            const float a = tap4_ref<true>(r.wa, mine[r.oa], mine[r.oa + 1], mine[r.oa + 2], mine[r.oa + 3]);
            const float b = tap4_ref<true>(r.wb, mine[r.ob], mine[r.ob + 1], mine[r.ob + 2], mine[r.ob + 3]);
            *op = lerp2_ref(a, b, r.t);
        }
    }
    // ---- longitude wrap (pseudo_pad.cu:82-96): left pad <- last `pad` columns, right pad <- first `pad` columns.
    // At most 2*pad columns per row; recomputed rather than read back.
    for (int j = warp; j < 2 * pad; j += NWARPS) {
        const int x = j < pad ? j : wl - 2 * pad + j;                 // source logical column
        if (x < x0 || x >= x1) continue;
        const int X = j < pad ? pad + wl + j : j - pad;               // destination physical column
        const ColRecipe r = rec[x - x0];
        float v = tap4_ref<true>(r.wa, mine[r.oa], mine[r.oa + 1], mine[r.oa + 2], mine[r.oa + 3]);
        if (hr >= 0) {
            const float b = tap4_ref<true>(r.wb, mine[r.ob], mine[r.ob + 1], mine[r.ob + 2], mine[r.ob + 3]);
            v = lerp2_ref(v, b, r.t);
        }
        orow[(i64)X * C + lane] = v;
    }
}


// ------------------------------------------------------------------------------------------------ slice + pad, TMA-staged
// Second version of the fused gather: persistent, warp-specialised, TMA-pipelined.
// The first one (above, kept for widths that are not a multiple of 4) is one small CTA per tile: half of its
// instructions are 4-byte cp.async staging and every CTA pays the whole table-load -> stage -> compute -> store latency
// chain with little memory traffic in flight (55 % of the HBM roofline; ncu, profiles/r1a_*).  Here
//   * CTAs are persistent (2 per SM) and walk "super-tiles" = (tile row y, band g, column chunk); inside one the
//     geometry and the per-column recipes are loop invariants and the CTA streams all images x 32-channel chunks;
//   * warp 0 is the producer: the circular span of every channel row is staged by 1-D bulk TMA copies
//     (`cp.async.bulk`, 16-byte aligned, one lane per channel row, completion on an mbarrier) into a 3-stage ring,
//     so ~120 KB of loads are in flight per SM while the eight consumer warps work;
//   * consumers: compute with lanes along the tile's x (each thread owns ONE column: its recipe stays in registers),
//     taps read from s_in[c][o..o+3], result to s_out[x][c] (pitch 33); then store with lanes along channels,
//     s_out[x][lane] -> 128-byte channels-last segments.  Both sides of the transpose are bank-conflict free.
constexpr int SROW = 160;          // floats per staged channel row: span (<= 152) + alignment slack, multiple of 4
constexpr int SP_STAGES = 3;
constexpr int SP_THREADS = 32 + 256;
constexpr size_t SP_SMEM = (size_t)SP_STAGES * CB * SROW * 4 + 2 * XB_MAX * (CB + 1) * 4 + 64;

struct SuperGeom {
    int g, y, x0, x1, wl, sb, erow, hr, wsrc, base_al, span_al, n1;
};

__device__ __forceinline__ SuperGeom super_geometry(i64 s, const Bands &bands, const SlicePadParams &P, const int *__restrict__ stab,
                                                    const int *__restrict__ hband, const int *__restrict__ hrow,
                                                    const int *__restrict__ hcol)
{
    SuperGeom G;
    const int W = P.W, h = P.h, pad = P.pad;
    G.y = (int)(s / P.xtotal);
    const int r = (int)(s % P.xtotal);
    int g = 0;
    while (g + 1 < bands.npart && P.xoff[g + 1] <= r) g++;
    G.g = g;
    G.wl = bands.wl[g];
    G.x0 = (r - P.xoff[g]) * P.xb[g];
    G.x1 = min(G.x0 + P.xb[g], G.wl);
    G.hr = -1;
    if (G.y >= pad && G.y < pad + h) {
        G.sb = g;
        G.erow = g * h + (G.y - pad);
    } else {
        const int sd = G.y < pad ? 0 : 1, rr = G.y < pad ? G.y : G.y - pad - h;
        G.hr = (g * 2 + sd) * pad + rr;
        G.sb = hband[G.hr];
        G.erow = G.sb * h + hrow[G.hr];
    }
    G.wsrc = bands.wl[G.sb];
    const int *tab = stab + (i64)G.sb * W;
    int base = 0, span = W + 3;
    if (W > P.scap) {
        int first, last;
        if (G.hr < 0) {
            first = tab[G.x0];
            last = tab[G.x1 - 1];
        } else {
            first = tab[hcol[(i64)G.hr * W + G.x0]];
            const int q = hcol[(i64)G.hr * W + G.x1 - 1];
            const int q1 = (q + 1 == G.wsrc) ? 0 : q + 1;
            last = tab[q1];
        }
        base = wrap_mod(first - 1, W);
        span = circ_dist(base, wrap_mod(last + 2, W), W) + 1;
        if (span > P.scap) __trap();              // host sizing guarantees this cannot happen
    }
    G.base_al = base & ~3;
    G.span_al = (span + (base - G.base_al) + 3) & ~3;          // <= SROW
    G.n1 = min(G.span_al, W - G.base_al);                      // floats before the longitude wrap (multiple of 4)
    return G;
}

__global__ void __launch_bounds__(SP_THREADS, 2) slice_pad_tma_kernel(const float *__restrict__ erp, float *__restrict__ out,
                                                                   const int *__restrict__ stab, const float4 *__restrict__ swt,
                                                                   const int *__restrict__ hband, const int *__restrict__ hrow,
                                                                   const int *__restrict__ hcol, const float *__restrict__ htw,
                                                                   Bands bands, SlicePadParams P)
{
    extern __shared__ __align__(128) unsigned char sp_smem[];
    float *s_in = reinterpret_cast<float *>(sp_smem);                                              // [SP_STAGES][CB][SROW]
    float *s_out = s_in + SP_STAGES * CB * SROW;                                                   // [2][XB_MAX][CB + 1]
    uint64_t *full = reinterpret_cast<uint64_t *>(s_out + 2 * XB_MAX * (CB + 1));                  // [SP_STAGES]
    uint64_t *empty = full + SP_STAGES;                                                            // [SP_STAGES]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, pad = P.pad, C = P.C;
    const int OH = h + 2 * pad;
    const i64 n_super = (i64)OH * P.xtotal;

    if (threadIdx.x == 0) {
        for (int i = 0; i < SP_STAGES; i++) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 8);             // one arrival per consumer warp
        }
        mbar_fence_init();
    }
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    if (warp == 0) {
        // ===================================================================================== producer
        for (i64 s = blockIdx.x; s < n_super; s += gridDim.x) {
            const SuperGeom G = super_geometry(s, bands, P, stab, hband, hrow, hcol);
            const int n2 = G.span_al - G.n1;
            for (int n = 0; n < P.N; n++) {
                for (int cc = 0; cc < P.cchunks; cc++) {
                    const int c0 = cc * CB, nc = min(CB, C - c0);
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (lane == 0) mbar_expect_tx(&full[stage], (uint32_t)(nc * G.span_al * 4));
                    __syncwarp();
                    if (lane < nc) {
                        const float *sr = erp + (((i64)n * C + c0 + lane) * (i64)(h * npart) + G.erow) * W;
                        float *dr = s_in + ((size_t)stage * CB + lane) * SROW;
                        bulk_g2s(dr, sr + G.base_al, (uint32_t)(G.n1 * 4), &full[stage]);
                        if (n2 > 0) bulk_g2s(dr + G.n1, sr, (uint32_t)(n2 * 4), &full[stage]);
                    }
                    if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ========================================================================================= consumers (8 warps)
    const int cw = warp - 1;
    const int xi = (cw & 3) * 32 + lane;          // this thread's column inside the chunk
    const int cg = (cw >> 2) * (CB / 2);          // and its 16-channel group
    int buf = 0;
    for (i64 s = blockIdx.x; s < n_super; s += gridDim.x) {
        const SuperGeom G = super_geometry(s, bands, P, stab, hband, hrow, hcol);
        const int nx = G.x1 - G.x0;
        const int *tab = stab + (i64)G.sb * W;
        const float4 *wtab = swt + (i64)G.sb * W;
        float4 wa = make_float4(0.f, 0.f, 0.f, 0.f), wb = wa;
        int oa = 0, ob = 0;
        float t = 0.f;
        if (xi < nx) {
            const int x = G.x0 + xi;
            if (G.hr < 0) {
                wa = wtab[x];
                const int o = tab[x] - 1 - G.base_al;
                oa = o < 0 ? o + W : o;
            } else {
                const i64 e = (i64)G.hr * W + x;
                const int q = hcol[e];
                const int q1 = (q + 1 == G.wsrc) ? 0 : q + 1;
                wa = wtab[q];
                wb = wtab[q1];
                int o = tab[q] - 1 - G.base_al;
                oa = o < 0 ? o + W : o;
                o = tab[q1] - 1 - G.base_al;
                ob = o < 0 ? o + W : o;
                t = htw[e];
            }
        }
        for (int n = 0; n < P.N; n++) {
            const i64 plane = (i64)n * npart + G.g;
            for (int cc = 0; cc < P.cchunks; cc++) {
                const int c0 = cc * CB, nc = min(CB, C - c0);
                float *so = s_out + (size_t)buf * XB_MAX * (CB + 1);
                const float *si = s_in + (size_t)stage * CB * SROW;
                mbar_wait(&full[stage], phase);
                if (xi < nx) {
                    float *dst = so + xi * (CB + 1) + cg;
                    if (G.hr < 0) {
#pragma unroll
                        for (int c = 0; c < CB / 2; c++) {
                            const float *r = si + (cg + c) * SROW + oa;
                            dst[c] = tap4_ref<true>(wa, r[0], r[1], r[2], r[3]);
                        }
                    } else {
#pragma unroll 8
                        for (int c = 0; c < CB / 2; c++) {
                            const float *ra = si + (cg + c) * SROW + oa;
                            const float *rb = si + (cg + c) * SROW + ob;
                            const float a = tap4_ref<true>(wa, ra[0], ra[1], ra[2], ra[3]);
                            const float b = tap4_ref<true>(wb, rb[0], rb[1], rb[2], rb[3]);
                            dst[c] = lerp2_ref(a, b, t);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stage]);                 // this warp no longer reads the stage
                if (++stage == SP_STAGES) { stage = 0; phase ^= 1; }
                asm volatile("bar.sync 1, 256;" ::: "memory");             // s_out[buf] complete (consumer warps only)

                // store: lanes along channels; the longitude wrap (pseudo_pad.cu:82-96) re-uses the staged results:
                // left pad <- last `pad` columns, right pad <- first `pad` columns
                float *orow = out + ((plane * OH + G.y) * (i64)P.out_pitch) * C + c0;
                if (lane < nc) {
                    // lean main loop (this loop was 38 % of the first TMA version's instructions): running pointers only
                    float *op = orow + (i64)(pad + G.x0 + cw) * C + lane;
                    const float *sp = so + cw * (CB + 1) + lane;
                    const i64 ostep = (i64)8 * C;
                    int i = cw;
#pragma unroll 1
                    for (; i + 24 < nx; i += 32, op += 4 * ostep, sp += 32 * (CB + 1)) {
                        const float v0 = sp[0], v1 = sp[8 * (CB + 1)], v2 = sp[16 * (CB + 1)], v3 = sp[24 * (CB + 1)];
                        op[0] = v0; op[ostep] = v1; op[2 * ostep] = v2; op[3 * ostep] = v3;
                    }
                    for (; i < nx; i += 8, op += ostep, sp += 8 * (CB + 1)) *op = *sp;
                    // wrap columns: at most 2 * pad per row, only in the first / last chunk of the band
                    if (G.x0 < pad || G.x1 > G.wl - pad) {
                        for (int j = cw; j < 2 * pad; j += 8) {
                            const int x = j < pad ? j : G.wl - 2 * pad + j;                 // source logical column
                            if (x < G.x0 || x >= G.x1) continue;
                            const int X = j < pad ? pad + G.wl + j : j - pad;               // destination physical column
                            orow[(i64)X * C + lane] = so[(x - G.x0) * (CB + 1) + lane];
                        }
                    }
                    if (P.zero_invalid) {
                        // columns beyond the band (PseudoPad leaves them 0; pseudo_pad.cu:39-54), split over the row's chunks
                        const int chunks = (G.wl + P.xb[G.g] - 1) / P.xb[G.g];
                        const int zb = (P.out_pitch - G.wl - 2 * pad + chunks - 1) / chunks;
                        const int X0 = G.wl + 2 * pad + (G.x0 / P.xb[G.g]) * zb, X1 = min(X0 + zb, P.out_pitch);
                        for (int X = X0 + cw; X < X1; X += 8) orow[(i64)X * C + lane] = 0.f;
                    }
                }
                buf ^= 1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ slice + pad, third version
// The TMA version above tops out at 0.67-0.70 of the HBM rate with nothing saturated (ncu r1g: 47 % SM throughput, 28 %
// warps active): every iteration is wait -> compute with lanes along x -> bar.sync -> store with lanes along channels, and
// in the narrow polar bands six of its eight consumer warps idle.  This version follows the structure that took the uslice
// gather to 0.88:
//   * one CTA per super-tile (tile row, band, column chunk), 4 CTAs / SM, looping over (image, 32-channel chunk) with
//     double-buffered 16-byte cp.async staging of the circular ERP span (rows 16-byte aligned, like the TMA rows);
//   * a lane owns FOUR CONSECUTIVE CHANNELS of a pixel (lane = 4 pixels x 8 channel quads), so the resampled values leave
//     as one 128-bit channels-last store per pixel - four full 128-byte lines per warp instruction - with NO second pass
//     through shared memory; the channel rows are skewed by 4 floats per channel quad, which makes the 4-byte tap loads
//     of a warp (8 quads x 4 consecutive columns) hit 32 different banks;
//   * per-pixel recipes (tap offsets, cubic weights, halo weight) are built once per CTA in shared memory.
// Same expression shapes (tap4_ref<true>, lerp2_ref) -> bit-identical tiles.
constexpr int S3_PITCH = 192;                 // floats per staged channel row: SROW (160) + the largest skew (28), = 0 mod 8
constexpr int S3_BUF = CB * S3_PITCH;         // floats per staging buffer (24 KB)
constexpr int S3_THREADS = 256;

__device__ __forceinline__ int s3_row(int c) { return c * S3_PITCH + (((c >> 2) & 7) << 2); }

__global__ void __launch_bounds__(S3_THREADS, 4) slice_pad_v3_kernel(const float *__restrict__ erp, float *__restrict__ out,
                                                                   const int *__restrict__ stab, const float4 *__restrict__ swt,
                                                                   const int *__restrict__ hband, const int *__restrict__ hrow,
                                                                   const int *__restrict__ hcol, const float *__restrict__ htw,
                                                                   Bands bands, SlicePadParams P)
{
    extern __shared__ __align__(128) unsigned char s3_smem[];
    float *s_in = reinterpret_cast<float *>(s3_smem);                               // [2][CB][S3_PITCH]
    ColRecipe *rec = reinterpret_cast<ColRecipe *>(s3_smem + 2 * S3_BUF * 4);       // [XB_MAX]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, pad = P.pad, C = P.C;
    const int OH = h + 2 * pad;

    const SuperGeom G = super_geometry(blockIdx.x, bands, P, stab, hband, hrow, hcol);
    const int nx = G.x1 - G.x0;
    const int n2 = G.span_al - G.n1;

    // ---- per-column recipes (tap offsets relative to the aligned span base)
    if (threadIdx.x < nx) {
        const int *tab = stab + (i64)G.sb * W;
        const float4 *wtab = swt + (i64)G.sb * W;
        const int x = G.x0 + threadIdx.x;
        ColRecipe r;
        if (G.hr < 0) {
            r.wa = wtab[x];
            const int o = tab[x] - 1 - G.base_al;
            r.oa = o < 0 ? o + W : o;
            r.wb = r.wa; r.ob = r.oa; r.t = 0.f;
        } else {
            const i64 e = (i64)G.hr * W + x;
            const int q = hcol[e];
            const int q1 = (q + 1 == G.wsrc) ? 0 : q + 1;
            r.wa = wtab[q];
            r.wb = wtab[q1];
            int o = tab[q] - 1 - G.base_al;
            r.oa = o < 0 ? o + W : o;
            o = tab[q1] - 1 - G.base_al;
            r.ob = o < 0 ? o + W : o;
            r.t = htw[e];
        }
        r.pad_ = 0;
        rec[threadIdx.x] = r;
    }

    // ---- staging: thread -> (channel row = tid / 8, 16-byte chunks tid % 8, +8, ...)
    const int sc = threadIdx.x >> 3, sq = threadIdx.x & 7;
    const int span4 = G.span_al >> 2, n14 = G.n1 >> 2;
    const i64 cstride = (i64)h * npart * W;
    const int iters = P.N * P.cchunks;
    auto stage = [&](int it, float *buf) {
        const int n = it / P.cchunks, c0 = (it - n * P.cchunks) * CB;
        if (sc < min(CB, C - c0)) {
            const float *sr = erp + (((i64)n * C + c0 + sc) * (i64)(h * npart) + G.erow) * W;
            float *dr = buf + s3_row(sc);
            for (int j = sq; j < span4; j += 8)
                cp_async16(dr + 4 * j, j < n14 ? sr + G.base_al + 4 * j : sr + 4 * (j - n14));
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, s_in);

    const int xsub = lane >> 3, cq = lane & 7;
    const int crow = s3_row(cq * 4);              // this lane's four channel rows: crow + m * S3_PITCH
    for (int it = 0; it < iters; it++) {
        const float *buf = s_in + (it & 1) * S3_BUF;
        if (it + 1 < iters) {
            stage(it + 1, s_in + ((it + 1) & 1) * S3_BUF);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();                           // staged data (and, the first time, the recipes) visible
        const int n = it / P.cchunks, c0 = (it - n * P.cchunks) * CB;
        const int c = c0 + cq * 4;
        if (c < C) {                               // C % 4 == 0: a lane's four channels are all valid or all not
            float *orow = out + ((((i64)n * npart + G.g) * OH + G.y) * (i64)P.out_pitch) * C + c;
            const float *mine = buf + crow;
            for (int xi = xsub + 4 * warp; xi < nx; xi += 32) {
                const ColRecipe r = rec[xi];
                float4 v;
                {
                    const float *t0 = mine + r.oa;
                    v.x = tap4_ref<true>(r.wa, t0[0], t0[1], t0[2], t0[3]);
                    v.y = tap4_ref<true>(r.wa, t0[S3_PITCH], t0[S3_PITCH + 1], t0[S3_PITCH + 2], t0[S3_PITCH + 3]);
                    v.z = tap4_ref<true>(r.wa, t0[2 * S3_PITCH], t0[2 * S3_PITCH + 1], t0[2 * S3_PITCH + 2], t0[2 * S3_PITCH + 3]);
                    v.w = tap4_ref<true>(r.wa, t0[3 * S3_PITCH], t0[3 * S3_PITCH + 1], t0[3 * S3_PITCH + 2], t0[3 * S3_PITCH + 3]);
                }
                if (G.hr >= 0) {
                    const float *t1 = mine + r.ob;
                    const float bx = tap4_ref<true>(r.wb, t1[0], t1[1], t1[2], t1[3]);
                    const float by = tap4_ref<true>(r.wb, t1[S3_PITCH], t1[S3_PITCH + 1], t1[S3_PITCH + 2], t1[S3_PITCH + 3]);
                    const float bz = tap4_ref<true>(r.wb, t1[2 * S3_PITCH], t1[2 * S3_PITCH + 1], t1[2 * S3_PITCH + 2], t1[2 * S3_PITCH + 3]);
                    const float bw = tap4_ref<true>(r.wb, t1[3 * S3_PITCH], t1[3 * S3_PITCH + 1], t1[3 * S3_PITCH + 2], t1[3 * S3_PITCH + 3]);
                    v.x = lerp2_ref(v.x, bx, r.t); v.y = lerp2_ref(v.y, by, r.t);
                    v.z = lerp2_ref(v.z, bz, r.t); v.w = lerp2_ref(v.w, bw, r.t);
                }
                const int x = G.x0 + xi;
                *reinterpret_cast<float4 *>(orow + (i64)(pad + x) * C) = v;
                // longitude wrap (pseudo_pad.cu:82-96): left pad <- last `pad` columns, right pad <- first `pad` columns
                if (x < pad) *reinterpret_cast<float4 *>(orow + (i64)(pad + G.wl + x) * C) = v;
                if (x >= G.wl - pad) *reinterpret_cast<float4 *>(orow + (i64)(x - (G.wl - pad)) * C) = v;
            }
            if (P.zero_invalid) {
                // columns beyond the band (PseudoPad leaves them 0; pseudo_pad.cu:39-54), split over the row's chunks
                const int chunks = (G.wl + P.xb[G.g] - 1) / P.xb[G.g];
                const int zb = (P.out_pitch - G.wl - 2 * pad + chunks - 1) / chunks;
                const int X0 = G.wl + 2 * pad + (G.x0 / P.xb[G.g]) * zb, X1 = min(X0 + zb, P.out_pitch);
                for (int X = X0 + xsub + 4 * warp; X < X1; X += 32)
                    *reinterpret_cast<float4 *>(orow + (i64)X * C) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
        __syncthreads();                           // everyone is done with this buffer before it is restaged
    }
}

// ------------------------------------------------------------------------------------------------ uslice
struct UsliceParams {
    int N, C, h, W;
    int in_rows, in_pitch, in_y0, in_x0;
    int xb, xchunks, cchunks, scap;
};

constexpr int UXB = 128;          // ERP columns per CTA
constexpr int UXJ = UXB / 32;     // columns per lane

__global__ void __launch_bounds__(NTHREADS) uslice_nhwc_kernel(const float *__restrict__ tiles, float *__restrict__ erp,
                                                               const int *__restrict__ utab, const float4 *__restrict__ uwt,
                                                               Bands bands, UsliceParams P)
{
    extern __shared__ float smu[];                // [span][CB + 1]
    constexpr int SP = CB + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, C = P.C;

    const int xc = blockIdx.x;
    const int cc = blockIdx.y % (unsigned)P.cchunks;
    const int y = blockIdx.y / (unsigned)P.cchunks;
    const i64 plane = blockIdx.z;
    const int g = blockIdx.z % (unsigned)npart;
    const i64 n = blockIdx.z / (unsigned)npart;
    const int wl = bands.wl[g];
    const int c0 = cc * CB;
    const int nc = min(CB, C - c0);
    const int X0 = xc * UXB, X1 = min(X0 + UXB, W);
    const int *tab = utab + (i64)g * W;
    const float4 *wtab = uwt + (i64)g * W;

    int base = 0, span = wl + 3;                  // small bands: whole row + the three wrapped columns
    if (wl > P.scap) {
        base = wrap_mod(tab[X0] - 1, wl);
        span = circ_dist(base, wrap_mod(tab[X1 - 1] + 2, wl), wl) + 1;
        if (span > P.scap) __trap();
    }

    // ---- stage: sm[i][c] = tile[plane][in_y0 + y][in_x0 + (base + i) mod wl][c0 + c], lanes along c
    const float *srow = tiles + ((plane * P.in_rows + P.in_y0 + y) * (i64)P.in_pitch + P.in_x0) * C + c0;
    if (lane < nc) {
        const int n1 = min(span, wl - base);
        const float *sp0 = srow + (i64)base * C + lane;
        float *dp = smu + lane;
        for (int i = warp; i < n1; i += NWARPS) cp_async4(dp + i * SP, sp0 + (i64)i * C);
        for (int i = n1 + warp; i < span; i += NWARPS) cp_async4(dp + i * SP, sp0 + (i64)(i - wl) * C);
    }

    // ---- this lane's columns: tap offsets and weights stay in registers for every channel
    int o[UXJ];
    float4 w[UXJ];
#pragma unroll
    for (int j = 0; j < UXJ; j++) {
        const int X = X0 + lane + 32 * j;
        o[j] = 0;
        w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (X < X1) {
            int t = tab[X] - 1 - base;
            o[j] = (t < 0 ? t + wl : t) * SP;
            w[j] = wtab[X];
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- produce: lanes along longitude, each warp owns channels warp, warp + 8, ...
    float *dst = erp + ((n * C + c0 + warp) * (i64)(h * npart) + (i64)g * h + y) * W + X0 + lane;
    const i64 cstep = (i64)NWARPS * h * npart * W;
    for (int c = warp; c < nc; c += NWARPS, dst += cstep) {
#pragma unroll
        for (int j = 0; j < UXJ; j++) {
            if (X0 + lane + 32 * j < X1) {
                const float *sp = smu + o[j] + c;
                dst[32 * j] = tap4_ref<false>(w[j], sp[0], sp[SP], sp[2 * SP], sp[3 * SP]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ uslice, second version
// The first version (above, kept for channel counts that are not a multiple of 4) is ISSUE-bound: 80 % SM throughput at
// 0.70 of the HBM rate (ncu, profiles/r1a_uslice_nhwc_kernel.json) - one 4-byte cp.async, four LDS.32 and one STG per output
// element.  Here
//   * staging moves 16 bytes per cp.async (8 threads cover the 128-byte channels-last segment of one source column) into
//     rows of exactly 128 bytes whose 16-byte chunks are XOR-swizzled with the row index, so that
//   * a thread produces FOUR channels of one ERP column from four LDS.128 (one per tap; a quarter-warp reads eight
//     consecutive rows -> eight different chunks, conflict free) and stores them with lanes along longitude;
//   * the CTA keeps its column recipes (16 tap addresses + 16 weights per thread) in registers and loops over the
//     32-channel chunks of C with double-buffered staging, so the next chunk's loads are in flight while this one is
//     resampled and the six 128-byte pieces of a 768-byte tile pixel are fetched back to back (DRAM page locality).
// ~6.5 thread-instructions per output element instead of ~11; bit-identical results (same tap4_ref expression).
constexpr int U2_ROWS = UXB + 8 + 4;          // staged source columns per buffer: scap (136) + 3 taps, rounded up

__global__ void __launch_bounds__(NTHREADS, 4) uslice_nhwc_v2_kernel(const float *__restrict__ tiles, float *__restrict__ erp,
                                                                  const int *__restrict__ utab, const float4 *__restrict__ uwt,
                                                                  Bands bands, UsliceParams P)
{
    __shared__ __align__(128) float smv[2][U2_ROWS * CB];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, C = P.C;

    const int xc = blockIdx.x;
    const int y = blockIdx.y;
    const i64 plane = blockIdx.z;
    const int g = blockIdx.z % (unsigned)npart;
    const i64 n = blockIdx.z / (unsigned)npart;
    const int wl = bands.wl[g];
    const int X0 = xc * UXB, X1 = min(X0 + UXB, W);
    const int *tab = utab + (i64)g * W;
    const float4 *wtab = uwt + (i64)g * W;

    int base = 0, span = wl + 3;                  // small bands: whole row + the three wrapped columns
    if (wl > P.scap) {
        base = wrap_mod(tab[X0] - 1, wl);
        span = circ_dist(base, wrap_mod(tab[X1 - 1] + 2, wl), wl) + 1;
        if (span > P.scap) __trap();
    }
    const int n1 = min(span, wl - base);          // staged columns before the longitude wrap

    // ---- staging: thread -> (source column i = tid / 8 (+32 per pass), 16-byte chunk q = tid % 8)
    const float *srow = tiles + ((plane * P.in_rows + P.in_y0 + y) * (i64)P.in_pitch + P.in_x0) * C;
    const int sq = threadIdx.x & 7, si0 = threadIdx.x >> 3;
    auto stage = [&](int cc, float *buf) {
        const int c0 = cc * CB;
        if (sq * 4 < min(CB, C - c0)) {
            const float *sp = srow + c0 + sq * 4;
            for (int i = si0; i < span; i += NTHREADS / 8) {
                const int col = i < n1 ? base + i : base + i - wl;
                cp_async16(buf + i * CB + ((sq ^ (i & 7)) << 2), sp + (i64)col * C);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    stage(0, smv[0]);

    // ---- this lane's columns: swizzled tap addresses and weights stay in registers for every channel chunk
    int a[UXJ][4];
    float4 w[UXJ];
#pragma unroll
    for (int j = 0; j < UXJ; j++) {
        const int X = X0 + lane + 32 * j;
        w[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        int o = 0;
        if (X < X1) {
            const int t = tab[X] - 1 - base;
            o = t < 0 ? t + wl : t;
            w[j] = wtab[X];
        }
#pragma unroll
        for (int k = 0; k < 4; k++) a[j][k] = (o + k) * CB + ((warp ^ ((o + k) & 7)) << 2);
    }

    const i64 cstride = (i64)h * npart * W;       // floats between channel planes of the ERP image
    float *drow = erp + (n * C * (i64)(h * npart) + (i64)g * h + y) * W + X0 + lane;
    for (int cc = 0; cc < P.cchunks; cc++) {
        const float *buf = smv[cc & 1];
        if (cc + 1 < P.cchunks) {
            stage(cc + 1, smv[(cc + 1) & 1]);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int c = cc * CB + warp * 4;         // this warp's four channels
        if (c < C) {
            float *dst = drow + c * cstride;
#pragma unroll
            for (int j = 0; j < UXJ; j++) {
                if (X0 + lane + 32 * j < X1) {
                    const float4 v0 = *reinterpret_cast<const float4 *>(buf + a[j][0]);
                    const float4 v1 = *reinterpret_cast<const float4 *>(buf + a[j][1]);
                    const float4 v2 = *reinterpret_cast<const float4 *>(buf + a[j][2]);
                    const float4 v3 = *reinterpret_cast<const float4 *>(buf + a[j][3]);
                    dst[32 * j] = tap4_ref<false>(w[j], v0.x, v1.x, v2.x, v3.x);
                    dst[32 * j + cstride] = tap4_ref<false>(w[j], v0.y, v1.y, v2.y, v3.y);
                    dst[32 * j + 2 * cstride] = tap4_ref<false>(w[j], v0.z, v1.z, v2.z, v3.z);
                    dst[32 * j + 3 * cstride] = tap4_ref<false>(w[j], v0.w, v1.w, v2.w, v3.w);
                }
            }
        }
        __syncthreads();                          // everyone is done with smv[cc & 1] before it is restaged
    }
}

}  // namespace

extern "C" {

int pcx_slice_pad_nhwc(const float *d_in, float *d_out, int N, int C, int H, int W, int npart, int pad, const int *wl,
                       const int *d_src, const float *d_wt, const int *d_band, const int *d_row, const int *d_col,
                       const float *d_tw, int out_pitch, int zero_invalid, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_in && d_out && d_src && d_wt, "null pointer");
    PCX_REQUIRE(pad == 0 || (d_band && d_row && d_col && d_tw), "halo tables are required when pad > 0");
    PCX_REQUIRE(N > 0 && C > 0 && W >= 4 && pad >= 0 && pad < 10, "bad shape N=%d C=%d W=%d pad=%d", N, C, W, pad);
    PCX_REQUIRE(H > 0 && H % npart == 0, "height %d is not a multiple of npart %d (math_cuda.cu:225)", H, npart);
    PCX_REQUIRE(out_pitch >= W + 2 * pad, "out_pitch %d < %d", out_pitch, W + 2 * pad);
    int wmin = W;
    for (int i = 0; i < npart; i++) {
        PCX_REQUIRE(wl[i] >= 4 && wl[i] >= 2 * pad && wl[i] <= W, "band %d width %d outside [max(4, 2 pad), %d]", i, wl[i], W);
        if (wl[i] < wmin) wmin = wl[i];
    }
    SlicePadParams P;
    P.N = N; P.C = C; P.h = H / npart; P.W = W; P.pad = pad; P.out_pitch = out_pitch; P.zero_invalid = zero_invalid;
    P.scap = 152;
    // ERP columns spanned by xb tile columns of band g: (xb-1) W/wl[g] for the columns themselves, up to 2 W/wl_src
    // for the second halo tap (bounded with the narrowest band), plus the cubic taps at both ends
    P.xchunks = 1;
    for (int i = 0; i < PCX_MAX_PART; i++) P.xb[i] = 8;
    for (int i = 0; i < npart; i++) {
        int xb = XB_MAX;
        if (W > P.scap) {
            const double ratio = (double)W / wl[i], worst = (double)W / wmin;
            while (xb > 8 && (xb - 1) * ratio + 2.0 * worst + 8.0 > P.scap) xb -= 8;
            PCX_REQUIRE((xb - 1) * ratio + 2.0 * worst + 8.0 <= P.scap, "band %d is too narrow (%d of %d columns) for the staged gather", i, wl[i], W);
        }
        P.xb[i] = xb;
        const int chunks = (wl[i] + xb - 1) / xb;
        if (chunks > P.xchunks) P.xchunks = chunks;
    }
    P.cchunks = (C + CB - 1) / CB;
    PCX_REQUIRE((i64)P.cchunks * (P.h + 2 * pad) <= 65535 && (i64)N * npart <= 65535, "grid too large");
    const dim3 grid(P.xchunks, P.cchunks * (P.h + 2 * pad), N * npart);
    const size_t smem = XB_MAX * sizeof(ColRecipe) + (size_t)CB * ((P.scap + 3) | 1) * sizeof(float);
    static const bool force_v1 = getenv("PCX_SLICE_V1") != nullptr || (getenv("PCX_SLICE_IMPL") && getenv("PCX_SLICE_IMPL")[0] == 'v');
    P.xoff[0] = 0;
    for (int i = 0; i < PCX_MAX_PART; i++) P.xoff[i + 1] = P.xoff[i] + (i < npart ? (wl[i] + P.xb[i] - 1) / P.xb[i] : 0);
    P.xtotal = P.xoff[npart];
    // PCX_SLICE_IMPL = tma | v1 forces one of the older kernels (A/B timing, tools/hbm_ops_bench.py)
    static const char *impl_env = getenv("PCX_SLICE_IMPL");
    static const bool force_tma = impl_env && impl_env[0] == 't';
    const bool aligned = W % 4 == 0 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0;
    if (aligned && C % 4 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && !force_v1 && !force_tma) {
        constexpr int smem3 = 2 * S3_BUF * 4 + XB_MAX * (int)sizeof(ColRecipe);
        static PcxDeviceOnce once3;
        PCX_ONCE_PER_DEVICE(once3) PCX_CUDA(cudaFuncSetAttribute(slice_pad_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3));
        const i64 n_super = (i64)(P.h + 2 * pad) * P.xtotal;
        PCX_REQUIRE(n_super < (1LL << 31), "grid too large");
        slice_pad_v3_kernel<<<(unsigned)n_super, S3_THREADS, smem3, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, d_band,
                                                                                          d_row, d_col, d_tw, b, P);
    } else if (aligned && !force_v1) {
        static PcxDeviceOnce once;
        PCX_ONCE_PER_DEVICE(once) PCX_CUDA(cudaFuncSetAttribute(slice_pad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SP_SMEM));
        const i64 n_super = (i64)(P.h + 2 * pad) * P.xtotal;
        const i64 ctas = 2LL * pcx_sm_count();
        slice_pad_tma_kernel<<<(unsigned)(n_super < ctas ? n_super : ctas), SP_THREADS, SP_SMEM, (cudaStream_t)stream>>>(
            d_in, d_out, d_src, (const float4 *)d_wt, d_band, d_row, d_col, d_tw, b, P);
    } else
        slice_pad_nhwc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, d_band,
                                                                                         d_row, d_col, d_tw, b, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_uslice_nhwc(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int in_rows, int in_pitch,
                    int in_y0, int in_x0, const int *wl, const int *d_src, const float *d_wt, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_in && d_out && d_src && d_wt, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W >= 4, "bad shape N=%d C=%d h=%d W=%d", N, C, h, W);
    PCX_REQUIRE(in_y0 >= 0 && in_x0 >= 0 && in_y0 + h <= in_rows && in_x0 + W <= in_pitch + 0, "tile window outside the input plane");
    for (int i = 0; i < npart; i++) PCX_REQUIRE(wl[i] >= 4 && wl[i] <= W, "band %d width %d outside [4,%d]", i, wl[i], W);
    UsliceParams P;
    P.N = N; P.C = C; P.h = h; P.W = W;
    P.in_rows = in_rows; P.in_pitch = in_pitch; P.in_y0 = in_y0; P.in_x0 = in_x0;
    P.xb = UXB;
    P.scap = UXB + 8;                             // wl <= W: at most one source column per destination column, + taps
    P.xchunks = (W + P.xb - 1) / P.xb;
    P.cchunks = (C + CB - 1) / CB;
    PCX_REQUIRE((i64)P.cchunks * h <= 65535 && (i64)N * npart <= 65535, "grid too large");
    static const bool force_v1 = getenv("PCX_USLICE_V1") != nullptr;
    if (C % 4 == 0 && (reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && !force_v1) {
        uslice_nhwc_v2_kernel<<<dim3(P.xchunks, h, N * npart), NTHREADS, 0, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, b, P);
        PCX_LAUNCHED();
        return PCX_OK;
    }
    const dim3 grid(P.xchunks, P.cchunks * h, N * npart);
    const size_t smem = (size_t)(P.scap + 3) * (CB + 1) * sizeof(float);
    uslice_nhwc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, b, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
