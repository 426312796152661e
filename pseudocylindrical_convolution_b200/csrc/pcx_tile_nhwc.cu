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

// ------------------------------------------------------------------------------------------------ slice + pad
struct SlicePadParams {
    int N, C, h, W, pad, out_pitch;
    int xb;            // logical tile columns per CTA
    int xchunks;       // column chunks per row (covers W)
    int cchunks;
    int scap;          // shared-memory row capacity in floats (odd pitch = scap | 1)
    int zero_invalid;  // write zeros to the columns >= wl + 2 pad
};

__global__ void __launch_bounds__(NTHREADS) slice_pad_nhwc_kernel(const float *__restrict__ erp, float *__restrict__ out,
                                                                  const int *__restrict__ stab, const float4 *__restrict__ swt,
                                                                  const int *__restrict__ hband, const int *__restrict__ hrow,
                                                                  const int *__restrict__ hcol, const float *__restrict__ htw,
                                                                  Bands bands, SlicePadParams P)
{
    extern __shared__ float sm[];
    const int spitch = P.scap | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, pad = P.pad, C = P.C;
    const int OH = h + 2 * pad;

    i64 bid = blockIdx.x;
    const int xc = (int)(bid % P.xchunks); bid /= P.xchunks;
    const int cc = (int)(bid % P.cchunks); bid /= P.cchunks;
    const int y = (int)(bid % OH); bid /= OH;
    const i64 plane = bid;
    const int g = (int)(plane % npart);
    const i64 n = plane / npart;
    const int wl = bands.wl[g];
    const int c0 = cc * CB;
    const int nc = min(CB, C - c0);
    const int x0 = xc * P.xb;
    const int x1 = min(x0 + P.xb, wl);            // logical columns [x0, x1) of the band

    float *orow = out + ((plane * OH + y) * (i64)P.out_pitch) * C + c0;

    // ---- columns beyond the band: zeros only (PseudoPad leaves them 0; pseudo_pad.cu:39-54)
    if (P.zero_invalid) {
        // physical columns handled by this chunk that lie at or beyond wl + 2 pad
        int X0 = x0 + 2 * pad, X1 = min(x0 + P.xb + 2 * pad, P.out_pitch);
        if (xc == P.xchunks - 1) X1 = P.out_pitch;
        if (X0 < wl + 2 * pad) X0 = wl + 2 * pad;
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

    // ---- circular span of ERP columns needed by logical columns [x0, x1)
    int base = 0, span = W;
    if (W > P.scap) {
        int xa = x0, xz = x1 - 1;
        int first, last;
        if (hr < 0) {
            first = tab[xa];
            last = tab[xz];
        } else {
            first = tab[hcol[(i64)hr * W + xa]];
            int q = hcol[(i64)hr * W + xz];
            int q1 = (q + 1 == wsrc) ? 0 : q + 1;
            last = tab[q1];
        }
        base = wrap_mod(first - 1, W);
        span = circ_dist(base, wrap_mod(last + 2, W), W) + 1;
        if (span > P.scap) __trap();              // host sizing guarantees this cannot happen
    }

    // ---- stage: sm[c][i] = erp[n][c0+c][erow][(base + i) mod W], lanes along i
    const float *src = erp + ((n * C + c0) * (i64)(h * npart) + erow) * W;
    const i64 cstride = (i64)h * npart * W;
    for (int c = warp; c < nc; c += NWARPS) {
        const float *sr = src + c * cstride;
        float *dr = sm + c * spitch;
        for (int i = lane; i < span; i += 32) {
            int col = base + i;
            if (col >= W) col -= W;
            dr[i] = __ldg(sr + col);
        }
    }
    __syncthreads();

    // resampled value of the source band at its column xs, for this lane's channel
    const float *mine = sm + lane * spitch;
    auto sliced = [&](int xs) -> float {
        const int p = tab[xs];
        const float4 w = wtab[xs];
        int o = p - 1 - base;
        if (o < 0) o += W;
        if (W > P.scap) {
            return tap4_ref<true>(w, mine[o], mine[o + 1], mine[o + 2], mine[o + 3]);
        } else {                                  // whole row staged: wrap each tap
            int o1 = o + 1 >= W ? o + 1 - W : o + 1;
            int o2 = o + 2 >= W ? o + 2 - W : o + 2;
            int o3 = o + 3 >= W ? o + 3 - W : o + 3;
            return tap4_ref<true>(w, mine[o], mine[o1], mine[o2], mine[o3]);
        }
    };

    for (int x = x0 + warp; x < x1; x += NWARPS) {
        float v = 0.f;
        if (lane < nc) {
            if (hr < 0) {
                v = sliced(x);
            } else {
                const i64 e = (i64)hr * W + x;
                const int q = hcol[e];
                const int q1 = (q + 1 == wsrc) ? 0 : q + 1;
                v = lerp2_ref(sliced(q), sliced(q1), htw[e]);
            }
            orow[(i64)(pad + x) * C + lane] = v;
            // longitude wrap (pseudo_pad.cu:82-96): left pad <- last `pad` columns, right pad <- first `pad` columns
            if (x < pad) orow[(i64)(pad + wl + x) * C + lane] = v;
            if (x >= wl - pad) orow[(i64)(x - (wl - pad)) * C + lane] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ uslice
struct UsliceParams {
    int N, C, h, W;
    int in_rows, in_pitch, in_y0, in_x0;
    int xb, xchunks, cchunks, scap;
};

__global__ void __launch_bounds__(NTHREADS) uslice_nhwc_kernel(const float *__restrict__ tiles, float *__restrict__ erp,
                                                               const int *__restrict__ utab, const float4 *__restrict__ uwt,
                                                               Bands bands, UsliceParams P)
{
    extern __shared__ float sm[];                 // [span][CB + 1]
    constexpr int SP = CB + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int npart = bands.npart;
    const int W = P.W, h = P.h, C = P.C;

    i64 bid = blockIdx.x;
    const int xc = (int)(bid % P.xchunks); bid /= P.xchunks;
    const int cc = (int)(bid % P.cchunks); bid /= P.cchunks;
    const int y = (int)(bid % h); bid /= h;
    const i64 plane = bid;
    const int g = (int)(plane % npart);
    const i64 n = plane / npart;
    const int wl = bands.wl[g];
    const int c0 = cc * CB;
    const int nc = min(CB, C - c0);
    const int X0 = xc * P.xb, X1 = min(X0 + P.xb, W);
    const int *tab = utab + (i64)g * W;
    const float4 *wtab = uwt + (i64)g * W;

    int base = 0, span = wl;
    if (wl > P.scap) {
        base = wrap_mod(tab[X0] - 1, wl);
        span = circ_dist(base, wrap_mod(tab[X1 - 1] + 2, wl), wl) + 1;
        if (span > P.scap) __trap();
    }
    const bool whole = !(wl > P.scap);

    // ---- stage: sm[i][c] = tile[plane][in_y0 + y][in_x0 + (base + i) mod wl][c0 + c], lanes along c
    const float *srow = tiles + ((plane * P.in_rows + P.in_y0 + y) * (i64)P.in_pitch + P.in_x0) * C + c0;
    for (int i = warp; i < span; i += NWARPS) {
        int col = base + i;
        if (col >= wl) col -= wl;
        if (lane < nc) sm[i * SP + lane] = __ldg(srow + (i64)col * C + lane);
    }
    __syncthreads();

    // ---- produce: lanes along longitude, each warp owns channels warp, warp + 8, ...
    float *dst = erp + ((n * C + c0) * (i64)(h * npart) + (i64)g * h + y) * W;
    const i64 cstride = (i64)h * npart * W;
    for (int X = X0 + lane; X < X1; X += 32) {
        const int p = tab[X];
        const float4 w = wtab[X];
        int o0 = p - 1 - base;
        if (o0 < 0) o0 += wl;
        int o1 = o0 + 1, o2 = o0 + 2, o3 = o0 + 3;
        if (whole) {
            if (o1 >= wl) o1 -= wl;
            if (o2 >= wl) o2 -= wl;
            if (o3 >= wl) o3 -= wl;
        }
        for (int c = warp; c < nc; c += NWARPS) {
            float v = tap4_ref<false>(w, sm[o0 * SP + c], sm[o1 * SP + c], sm[o2 * SP + c], sm[o3 * SP + c]);
            dst[c * cstride + X] = v;
        }
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
    P.scap = 320;
    // ERP columns spanned by xb tile columns: (xb-1) W/wl for the columns themselves, up to 2 W/wl_src for the
    // second halo tap, plus the cubic taps at both ends
    const double ratio = (double)W / wmin;
    P.xb = 64;
    while (P.xb > 1 && (P.xb - 1) * ratio + 2.0 * ratio + 8.0 > P.scap) P.xb >>= 1;
    if (W <= P.scap) P.xb = 64;                   // whole rows are staged: no span limit
    P.xchunks = (W + P.xb - 1) / P.xb;
    P.cchunks = (C + CB - 1) / CB;
    const i64 blocks = (i64)N * npart * (P.h + 2 * pad) * P.cchunks * P.xchunks;
    PCX_REQUIRE(blocks < (1ll << 31), "grid too large");
    const size_t smem = (size_t)CB * (P.scap | 1) * sizeof(float);
    slice_pad_nhwc_kernel<<<(unsigned)blocks, NTHREADS, smem, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, d_band,
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
    P.xb = 128;
    P.scap = P.xb + 8;                            // wl <= W: at most one source column per destination column, + taps
    P.xchunks = (W + P.xb - 1) / P.xb;
    P.cchunks = (C + CB - 1) / CB;
    const i64 blocks = (i64)N * npart * h * P.cchunks * P.xchunks;
    PCX_REQUIRE(blocks < (1ll << 31), "grid too large");
    const size_t smem = (size_t)P.scap * (CB + 1) * sizeof(float);
    uslice_nhwc_kernel<<<(unsigned)blocks, NTHREADS, smem, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, b, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
