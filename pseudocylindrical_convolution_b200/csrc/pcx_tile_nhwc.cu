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

// 4-byte asynchronous global -> shared copy (LDGSTS): no register round trip, the warp keeps issuing
__device__ __forceinline__ void cp_async4(float *dst_smem, const float *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ slice + pad
struct SlicePadParams {
    int N, C, h, W, pad, out_pitch;
    int xb[PCX_MAX_PART];   // logical tile columns per CTA, per band (narrow polar bands span more ERP columns per column)
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
    const dim3 grid(P.xchunks, P.cchunks * h, N * npart);
    const size_t smem = (size_t)(P.scap + 3) * (CB + 1) * sizeof(float);
    uslice_nhwc_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(d_in, d_out, d_src, (const float4 *)d_wt, b, P);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
