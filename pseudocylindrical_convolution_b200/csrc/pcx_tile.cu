// pcx_tile.cu - the pseudocylindrical tile pipeline: ERP <-> latitude bands, halo construction, fill,
// depth<->space.  HBM-bound kernels.
//
// Common structure ("row pipeline"): a CTA owns a strip of rows of one (image, band, channel) plane.
// Source rows are staged in shared memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier, one elected
// thread, STAGES rows in flight), every thread then produces 4 consecutive output columns from shared
// memory and writes them with one 128-bit streaming store.  Per-column gather tables (integer tap +
// Catmull-Rom weights) are band constants and stay in registers for the whole strip, so HBM sees exactly
// one read of the source row and one write of the destination row.
#include "pcx_common.cuh"

namespace {

constexpr int STAGES = 3;
constexpr int ROW_LEAD = 4;   // floats in front of a staged row (room for the wrap tap at column -1, keeps 16B alignment)
constexpr int ROW_TAIL = 4;   // floats behind it (wrap taps at columns n, n+1)

__host__ __device__ inline int staged_row_floats(int n) { return ROW_LEAD + ((n + 3) & ~3) + ROW_TAIL; }

struct RowStager {
    uint64_t *bars;
    float *rows;
    int row_floats;
    bool bulk;

    __device__ void init(unsigned char *smem, int max_row, bool use_bulk)
    {
        bars = reinterpret_cast<uint64_t *>(smem);
        rows = reinterpret_cast<float *>(smem + 64);
        row_floats = staged_row_floats(max_row);
        bulk = use_bulk;
        if (threadIdx.x == 0) {
            for (int s = 0; s < STAGES; s++) mbar_init(&bars[s], 1);
            mbar_fence_init();
        }
        __syncthreads();
    }
    __device__ float *row(int stage) { return rows + (size_t)stage * row_floats + ROW_LEAD; }
    // called by every thread; src must stay valid until wait()
    __device__ void issue(int stage, const float *src, int n)
    {
        float *dst = row(stage);
        if (bulk) {
            if (threadIdx.x == 0) {
                uint32_t bytes = (uint32_t)(((n + 3) & ~3) * sizeof(float));
                mbar_expect_tx(&bars[stage], bytes);
                bulk_g2s(dst, src, bytes, &bars[stage]);
            }
        } else {
            for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = __ldg(src + i);
        }
    }
    // after wait() returns the first n floats of row(stage) are visible to every thread, and the wrap taps
    // row[-1] = row[n-1], row[n] = row[0], row[n+1] = row[1] are in place.
    __device__ void wait(int stage, uint32_t parity, int n)
    {
        if (bulk) mbar_wait(&bars[stage], parity);
        else __syncthreads();
        float *r = row(stage);
        if (threadIdx.x == 0) r[-1] = r[n - 1];
        if (threadIdx.x == 32 % blockDim.x) { float a = r[0], b = r[1]; r[n] = a; r[n + 1] = b; }
        __syncthreads();
    }
};

inline size_t stager_smem(int max_row) { return 64 + (size_t)STAGES * staged_row_floats(max_row) * sizeof(float); }

// ------------------------------------------------------------------------------------------------ slice / uslice
// TO_TILES = true : sphere_slice_forward_kernel  (extension/sphere_slice_cuda.cu:87-116)
// TO_TILES = false: sphere_uslice_forward_kernel (extension/sphere_uslice_cuda.cu:73-99)
// CH = column chunks per thread (each 4 columns); blockDim.x * 4 * CH >= W.
template <int CH, bool TO_TILES, int MAXT>
__global__ void __launch_bounds__(MAXT) resample_rows_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                             const int *__restrict__ src_tab, const float4 *__restrict__ wt_tab,
                                                             Bands bands, int C, int h, int W, int pad, int rows_per_cta,
                                                             int strips, bool bulk, bool vec_store)
{
    extern __shared__ __align__(128) unsigned char smem[];
    RowStager st;
    st.init(smem, W, bulk);

    const int npart = bands.npart;
    i64 plane = blockIdx.x / strips;          // (n * npart + g) * C + c   in tile order
    int strip = blockIdx.x % strips;
    int c = (int)(plane % C);
    int g = (int)((plane / C) % npart);
    i64 n = plane / C / npart;
    const int wl = bands.wl[g];
    const int y0 = strip * rows_per_cta;
    const int nrows = min(rows_per_cta, h - y0);
    const i64 H = (i64)h * npart;
    const i64 tile_h = h + 2 * pad, tile_w = W + 2 * pad;

    // source / destination row 0 of this strip
    const float *src0;
    float *dst0;
    i64 src_pitch, dst_pitch;
    int n_src;           // samples per staged source row
    if (TO_TILES) {
        src0 = in + ((n * C + c) * H + (i64)g * h + y0) * W;
        src_pitch = W;
        n_src = W;
        dst0 = out + (plane * tile_h + y0 + pad) * tile_w + pad;
        dst_pitch = tile_w;
    } else {
        src0 = in + (plane * tile_h + y0 + pad) * tile_w + pad;
        src_pitch = tile_w;
        n_src = wl;
        dst0 = out + ((n * C + c) * H + (i64)g * h + y0) * W;
        dst_pitch = W;
    }

    // band-constant gather table for this thread's columns
    int tap[CH][4];
    float4 wt[CH][4];
#pragma unroll
    for (int ch = 0; ch < CH; ch++) {
        int x0 = (ch * blockDim.x + threadIdx.x) * 4;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int x = x0 + j;
            bool live = x < W && (!TO_TILES || x < wl);
            tap[ch][j] = live ? src_tab[(i64)g * W + x] : -1;
            wt[ch][j] = live ? wt_tab[(i64)g * W + x] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }

    for (int s = 0; s < STAGES && s < nrows; s++) st.issue(s, src0 + (i64)s * src_pitch, n_src);

    for (int r = 0; r < nrows; r++) {
        const int stage = r % STAGES;
        st.wait(stage, (r / STAGES) & 1, n_src);
        const float *row = st.row(stage);
        float *dst = dst0 + (i64)r * dst_pitch;
#pragma unroll
        for (int ch = 0; ch < CH; ch++) {
            int x0 = (ch * blockDim.x + threadIdx.x) * 4;
            if (x0 >= W) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                int p = tap[ch][j];
                v[j] = (p < 0) ? 0.f : tap4_ref<TO_TILES>(wt[ch][j], row[p - 1], row[p], row[p + 1], row[p + 2]);
            }
            if (vec_store) {
                st_cs_f4(dst + x0, make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++)
                    if (x0 + j < W) dst[x0 + j] = v[j];
            }
        }
        __syncthreads();                         // everyone is done with this stage
        if (r + STAGES < nrows) st.issue(stage, src0 + (i64)(r + STAGES) * src_pitch, n_src);
    }
}

// ------------------------------------------------------------------------------------------------ pad
// pseudo_pad_copy_forward_kernel + pseudo_pad_forward_kernel + pseudo_pad_circle_forward_kernel
// (extension/pseudo_pad.cu:39-96) in one pass.  CAUSAL selects the PseudoEntropyPad variant
// (extension/pseudo_entropy_pad_cuda.cu:39-105): pole rows 0, left pad 0, missing left sample 0.
// One CTA = a strip of OUTPUT rows of one (tile, channel) plane; VEC = store width in floats.
template <int VEC, bool CAUSAL>
__global__ void __launch_bounds__(512) pad_rows_kernel(const float *__restrict__ in, float *__restrict__ out, Bands bands,
                                                       const int *__restrict__ hband, const int *__restrict__ hrow,
                                                       const int *__restrict__ hcol, const float *__restrict__ htw,
                                                       int C, int h, int W, int pad, int out_pitch, int rows_per_cta,
                                                       int strips, bool bulk)
{
    extern __shared__ __align__(128) unsigned char smem[];
    RowStager st;
    st.init(smem, W, bulk);

    const int npart = bands.npart;
    i64 plane = blockIdx.x / strips;          // tile * C + c
    int strip = blockIdx.x % strips;
    int c = (int)(plane % C);
    i64 tile = plane / C;
    int g = (int)(tile % npart);
    i64 img = tile / npart;
    const int wl = bands.wl[g];
    const int out_h = h + 2 * pad;
    const int y0 = strip * rows_per_cta;
    const int nrows = min(rows_per_cta, out_h - y0);

    // source row feeding output row y: interior -> own plane, halo -> neighbour band (none at a causal pole)
    auto source = [&](int y, const float *&src, int &n, int &hr) {
        if (y >= pad && y < pad + h) {
            src = in + (plane * h + (y - pad)) * (i64)W;
            n = wl;
            hr = -1;
        } else {
            int s = y < pad ? 0 : 1;
            int r = y < pad ? y : y - pad - h;
            hr = (g * 2 + s) * pad + r;
            int pg = hband[hr];
            if (pg < 0) { src = nullptr; n = 0; return; }
            src = in + (((img * npart + pg) * C + c) * h + hrow[hr]) * (i64)W;
            n = bands.wl[pg];
        }
    };

    for (int s = 0; s < STAGES && s < nrows; s++) {
        const float *src; int n, hr;
        source(y0 + s, src, n, hr);
        if (src) st.issue(s, src, n);
    }

    uint32_t phase = 0;     // bit s = parity to wait for on stage s (pole rows skip issue and wait alike)
    for (int r = 0; r < nrows; r++) {
        const int stage = r % STAGES;
        const int y = y0 + r;
        const float *src; int n, hr;
        source(y, src, n, hr);
        if (src) {
            st.wait(stage, (phase >> stage) & 1u, n);
            phase ^= 1u << stage;
        }
        const float *row = st.row(stage);
        float *dst = out + (plane * out_h + y) * (i64)out_pitch;
        for (int x0 = threadIdx.x * VEC; x0 < out_pitch; x0 += blockDim.x * VEC) {
            float v[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                int x = x0 + j - pad;                    // column in the unpadded tile
                if (x < 0) x = CAUSAL ? -1 : x + wl;     // left pad: last `pad` valid columns (0 when causal)
                else if (x >= wl) { x -= wl; if (x >= pad) x = -1; }   // right pad: first `pad` valid columns
                float val = 0.f;
                if (x >= 0 && src) {
                    if (hr < 0) {
                        val = row[x];
                    } else {
                        i64 e = (i64)hr * W + x;
                        int q = hcol[e];
                        float a = (CAUSAL && q < 0) ? 0.f : row[q];
                        val = lerp2_ref(a, row[q + 1], htw[e]);
                    }
                }
                v[j] = val;
            }
            if (VEC == 4) st_cs_f4(dst + x0, make_float4(v[0], v[1], v[2], v[3]));
            else if (VEC == 2) st_cs_f2(dst + x0, make_float2(v[0], v[1]));
            else dst[x0] = v[0];
        }
        __syncthreads();
        if (r + STAGES < nrows) {
            const float *nsrc; int nn, nhr;
            source(y + STAGES, nsrc, nn, nhr);
            if (nsrc) st.issue(stage, nsrc, nn);
        }
    }
}

// Same operator without the shared-memory row pipeline: one WARP per output row, lanes stride over the row in vectors of VEC
// columns, everything read straight from global memory (a padded row is its source row shifted by `pad` columns, so the scalar
// loads of a warp are contiguous; the wrap columns and the 2-tap halo rows hit L1/L2).  No block-wide synchronisation at all:
// this is what lets the copy stream at HBM speed - the staged version above spends two barriers per 8 KB row.
template <bool CAUSAL>
__device__ __forceinline__ float pad_value(const float *__restrict__ src, int hr, int xo, int wl, int wsrc, int W, int pad,
                                           const int *__restrict__ hcol, const float *__restrict__ htw)
{
    int x = xo - pad;                        // column in the unpadded tile
    if (x < 0) x = CAUSAL ? -1 : x + wl;     // left pad: last `pad` valid columns (0 when causal)
    else if (x >= wl) { x -= wl; if (x >= pad) x = -1; }   // right pad: first `pad` valid columns
    if (x < 0 || src == nullptr) return 0.f;
    if (hr < 0) return __ldg(src + x);
    const i64 e = (i64)hr * W + x;
    const int q = hcol[e];
    const float a = (CAUSAL && q < 0) ? 0.f : __ldg(src + (q < 0 ? q + wsrc : q));
    const int q1 = q + 1 >= wsrc ? q + 1 - wsrc : q + 1;
    return lerp2_ref(a, __ldg(src + q1), htw[e]);
}

// The bulk of a row is written with 16-byte-aligned 128-bit stores whatever the pitch (2050-float rows of pad = 1 start on
// 8-byte boundaries): up to three leading and trailing columns are peeled off as scalar stores.
template <bool CAUSAL>
__global__ void __launch_bounds__(256) pad_direct_kernel(const float *__restrict__ in, float *__restrict__ out, Bands bands,
                                                         const int *__restrict__ hband, const int *__restrict__ hrow,
                                                         const int *__restrict__ hcol, const float *__restrict__ htw,
                                                         i64 nrows, int C, int h, int W, int pad, int out_pitch)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    const int npart = bands.npart, out_h = h + 2 * pad;
    for (i64 R = warp0; R < nrows; R += nwarps) {
        const i64 plane = R / out_h;
        const int y = (int)(R - plane * out_h);
        const int c = (int)(plane % C);
        const i64 tile = plane / C;
        const int g = (int)(tile % npart);
        const int wl = bands.wl[g];
        const float *src = nullptr;
        int hr = -1, wsrc = wl;
        if (y >= pad && y < pad + h) {
            src = in + (plane * h + (y - pad)) * (i64)W;
        } else {
            const int s = y < pad ? 0 : 1, r = y < pad ? y : y - pad - h;
            hr = (g * 2 + s) * pad + r;
            const int pg = hband[hr];
            if (pg >= 0) {
                src = in + ((((tile / npart) * npart + pg) * C + c) * h + hrow[hr]) * (i64)W;
                wsrc = bands.wl[pg];
            }
        }
        float *dst = out + R * (i64)out_pitch;
        const int lead = (int)(((16 - (reinterpret_cast<uintptr_t>(dst) & 15)) & 15) >> 2);       // floats before alignment
        const int nv = (out_pitch - lead) >> 2;                                                  // aligned vectors
        if (lane < lead) dst[lane] = pad_value<CAUSAL>(src, hr, lane, wl, wsrc, W, pad, hcol, htw);
        const int tail0 = lead + nv * 4;
        if (tail0 + lane < out_pitch && lane < 4) dst[tail0 + lane] = pad_value<CAUSAL>(src, hr, tail0 + lane, wl, wsrc, W, pad, hcol, htw);
        if (src != nullptr && hr < 0) {
            // interior row: vectors [ivB0, ivB1) are plain shifted copies, [ivB1, ivC) hold the right wrap, the rest is zero
            const int ivB0 = lead >= pad ? 0 : (pad - lead + 3) >> 2;
            int ivB1 = (wl + pad - lead) >> 2;                     // first vector with x0 + 3 - pad >= wl
            ivB1 = ivB1 < ivB0 ? ivB0 : (ivB1 > nv ? nv : ivB1);
            int ivC = (wl + 2 * pad - lead + 3) >> 2;              // first vector that lies wholly beyond the wrap columns
            ivC = ivC < ivB1 ? ivB1 : (ivC > nv ? nv : ivC);
            const float *sp = src + lead - pad;
            int iv = ivB0 + lane;
            for (; iv + 96 < ivB1; iv += 128) {                    // four vectors per lane, sixteen loads in flight
                float v[4][4];
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int j = 0; j < 4; j++) v[u][j] = __ldg(sp + (iv + 32 * u) * 4 + j);
#pragma unroll
                for (int u = 0; u < 4; u++) st_cs_f4(dst + lead + (iv + 32 * u) * 4, make_float4(v[u][0], v[u][1], v[u][2], v[u][3]));
            }
            for (; iv < ivB1; iv += 32) {
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = __ldg(sp + iv * 4 + j);
                st_cs_f4(dst + lead + iv * 4, make_float4(v[0], v[1], v[2], v[3]));
            }
            for (iv = lane; iv < ivB0; iv += 32) {                  // left wrap
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = pad_value<CAUSAL>(src, hr, lead + iv * 4 + j, wl, wsrc, W, pad, hcol, htw);
                st_cs_f4(dst + lead + iv * 4, make_float4(v[0], v[1], v[2], v[3]));
            }
            for (iv = ivB1 + lane; iv < ivC; iv += 32) {            // last valid columns + right wrap
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = pad_value<CAUSAL>(src, hr, lead + iv * 4 + j, wl, wsrc, W, pad, hcol, htw);
                st_cs_f4(dst + lead + iv * 4, make_float4(v[0], v[1], v[2], v[3]));
            }
            for (iv = ivC + lane; iv < nv; iv += 32) st_cs_f4(dst + lead + iv * 4, make_float4(0.f, 0.f, 0.f, 0.f));
        } else {
            for (int iv = lane; iv < nv; iv += 32) {                // halo rows (2-tap gathers) and pole rows (zeros)
                const int x0 = lead + iv * 4;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = pad_value<CAUSAL>(src, hr, x0 + j, wl, wsrc, W, pad, hcol, htw);
                st_cs_f4(dst + x0, make_float4(v[0], v[1], v[2], v[3]));
            }
        }
    }
}

// In-place halo refresh of a padded, pitched buffer: only halo rows and wrap columns are touched.
// Same values as pad_rows_kernel<*, false>.  One thread per written cell.
__global__ void halo_fill_kernel(float *__restrict__ buf, Bands bands, const int *__restrict__ hband,
                                 const int *__restrict__ hrow, const int *__restrict__ hcol,
                                 const float *__restrict__ htw, i64 planes, int C, int h, int W, int pad, int pitch)
{
    const int npart = bands.npart;
    const int out_h = h + 2 * pad;
    const int halo_cells = 2 * pad * (W + 2 * pad);
    const int side_cells = h * 2 * pad;
    const i64 per_plane = halo_cells + side_cells;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < planes * per_plane; idx += (i64)gridDim.x * blockDim.x) {
        i64 plane = idx / per_plane;
        int k = (int)(idx % per_plane);
        int c = (int)(plane % C);
        i64 tile = plane / C;
        int g = (int)(tile % npart);
        i64 img = tile / npart;
        const int wl = bands.wl[g];
        float *pl = buf + plane * out_h * (i64)pitch;
        if (k >= halo_cells) {              // wrap cell of an interior row
            k -= halo_cells;
            int y = pad + k / (2 * pad), j = k % (2 * pad);
            float *rowp = pl + (i64)y * pitch;
            if (j < pad) rowp[j] = rowp[pad + wl - pad + j];
            else rowp[pad + wl + (j - pad)] = rowp[pad + (j - pad)];
            continue;
        }
        int hy = k / (W + 2 * pad), xo = k % (W + 2 * pad);
        int y = hy < pad ? hy : h + hy;      // halo rows: [0,pad) and [pad+h, 2pad+h)
        int s = hy < pad ? 0 : 1, r = hy < pad ? hy : hy - pad;
        int hr = (g * 2 + s) * pad + r;
        int x = xo - pad;
        if (x < 0) x += wl;
        else if (x >= wl) { x -= wl; if (x >= pad) x = -1; }
        float val = 0.f;
        if (x >= 0) {
            int pg = hband[hr];
            const float *srow = buf + ((((img * npart + pg) * C + c) * out_h) + pad + hrow[hr]) * (i64)pitch + pad;
            i64 e = (i64)hr * W + x;
            int q = hcol[e];
            int q1 = (q + 1 == bands.wl[pg]) ? 0 : q + 1;
            val = lerp2_ref(srow[q], srow[q1], htw[e]);
        }
        pl[(i64)y * pitch + xo] = val;
    }
}

// pseudo_fill_forward_kernel (extension/pseudo_fill_cuda.cu:28-43): one warp per row, only the cells
// that change are written.
__global__ void fill_rows_kernel(float *__restrict__ data, Bands bands, i64 rows, int C, int Hh, int Ww, int pad, int trim, float fvalue)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const i64 nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 rowi = warp0; rowi < rows; rowi += nwarps) {
        int y = (int)(rowi % Hh);
        int g = (int)((rowi / Hh / C) % bands.npart);
        float *p = data + rowi * Ww;
        int lo = pad - trim, hi = pad + bands.wl[g] + trim;      // valid columns are [lo, hi)
        if (y < pad - trim || y >= Hh - pad + trim) { lo = Ww; hi = Ww; }   // whole row
        if (lo > Ww) lo = Ww;
        if (lo < 0) lo = 0;
        if (hi < lo) hi = lo;
        for (int x = lane; x < lo; x += 32) p[x] = fvalue;
        for (int x = hi + lane; x < Ww; x += 32) p[x] = fvalue;
    }
}

// dtow_forward_kernel / wtod_forward_kernel (extension/dtow_cuda.cu:38-75), output-driven so that the
// stores are coalesced.
__global__ void dtow_kernel(const float *__restrict__ in, float *__restrict__ out, i64 total, int C, int H, int W, int s, bool d2w)
{
    const int s2 = s * s;
    for (i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (i64)gridDim.x * blockDim.x) {
        if (d2w) {
            const int Co = C / s2; const i64 Ho = (i64)H * s, Wo = (i64)W * s;
            i64 px = idx % Wo, py = (idx / Wo) % Ho; i64 pc = (idx / Wo / Ho) % Co, n = idx / Wo / Ho / Co;
            i64 c = pc * s2 + (py % s) * s + (px % s);
            out[idx] = in[((n * C + c) * H + py / s) * W + px / s];
        } else {
            const i64 Co = (i64)C * s2; const i64 Ho = H / s, Wo = W / s;
            i64 px = idx % Wo, py = (idx / Wo) % Ho; i64 pc = (idx / Wo / Ho) % Co, n = idx / Wo / Ho / Co;
            i64 c = pc / s2; int rc = (int)(pc % s2);
            out[idx] = in[((n * C + c) * H + py * s + rc / s) * W + px * s + rc % s];
        }
    }
}

// stride-2 fast paths: 128-bit stores, 64/128-bit loads, 32-bit index arithmetic.
// d2w: out (N, C/4, 2H, 2W); four output columns come from two consecutive columns of two input channels.
__global__ void __launch_bounds__(256) dtow2_d2w_kernel(const float *__restrict__ in, float *__restrict__ out, unsigned nvec, int C,
                                                        int H, int W)
{
    const unsigned Co = C / 4, Ho = 2 * H, WV = (2 * W) / 4;
    const i64 chs = (i64)H * W;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        const unsigned row = i / WV, xv = i - row * WV;
        const unsigned py = row % Ho, pcn = row / Ho;
        const unsigned pc = pcn % Co, n = pcn / Co;
        const float *src = in + (((i64)n * C + pc * 4 + (py & 1) * 2) * H + (py >> 1)) * W + xv * 2;
        const float2 a = __ldcs(reinterpret_cast<const float2 *>(src));
        const float2 b = __ldcs(reinterpret_cast<const float2 *>(src + chs));
        st_cs_f4(out + (i64)i * 4, make_float4(a.x, b.x, a.y, b.y));
    }
}
// w2d: out (N, 4C, H/2, W/2); eight consecutive input columns of one row feed four columns of two output channels.
__global__ void __launch_bounds__(256) dtow2_w2d_kernel(const float *__restrict__ in, float *__restrict__ out, unsigned nvec, int C,
                                                        int H, int W)
{
    const unsigned Ho = H / 2, Wo = W / 2, WV = Wo / 4;
    const i64 ochs = (i64)Ho * Wo;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += gridDim.x * blockDim.x) {
        unsigned t = i / WV;
        const unsigned xv = i - t * WV;
        const unsigned r = t & 1; t >>= 1;
        const unsigned oy = t % Ho; t /= Ho;
        const unsigned c = t % (unsigned)C, n = t / (unsigned)C;
        const float *src = in + (((i64)n * C + c) * H + oy * 2 + r) * W + xv * 8;
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(src));
        const float4 b = __ldcs(reinterpret_cast<const float4 *>(src + 4));
        float *dst = out + (((i64)n * C * 4 + c * 4 + r * 2) * Ho + oy) * Wo + xv * 4;
        st_cs_f4(dst, make_float4(a.x, a.z, b.x, b.z));
        st_cs_f4(dst + ochs, make_float4(a.y, a.w, b.y, b.w));
    }
}

// ------------------------------------------------------------------------------------------------ launch helpers
struct StripPlan { int rows_per_cta, strips; };

// rows of one plane are split so that the grid has a few CTAs per SM, but a CTA keeps at least 4 rows so
// the band-constant tables in registers are amortised.
StripPlan plan_strips(i64 planes, int rows)
{
    i64 want = (i64)pcx_sm_count() * 6;
    int strips = 1;
    if (planes < want) strips = (int)((want + planes - 1) / planes);
    int max_strips = (rows + 3) / 4;
    if (strips > max_strips) strips = max_strips;
    if (strips < 1) strips = 1;
    int rpc = (rows + strips - 1) / strips;
    strips = (rows + rpc - 1) / rpc;
    return {rpc, strips};
}

template <typename K>
int allow_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e != cudaSuccess) { pcx_set_error("cudaFuncSetAttribute(%zu): %s", bytes, cudaGetErrorString(e)); return PCX_ECUDA; }
    }
    return PCX_OK;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <bool TO_TILES>
int launch_resample(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, const int *wl,
                    const int *d_src, const float *d_wt, int pad, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_in && d_out && d_src && d_wt, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W >= 4 && pad >= 0, "bad shape N=%d C=%d h=%d W=%d pad=%d", N, C, h, W, pad);
    PCX_REQUIRE(W <= 8192, "W=%d exceeds the 8192-column limit of the row pipeline", W);
    for (int i = 0; i < npart; i++) PCX_REQUIRE(wl[i] >= 4 && wl[i] <= W, "band %d width %d outside [4,%d]", i, wl[i], W);
    i64 planes = (i64)N * npart * C;
    StripPlan sp = plan_strips(planes, h);
    PCX_REQUIRE(planes * sp.strips < (1ll << 31), "grid too large");
    // source rows: ERP rows (pitch W) for slice, tile rows (pitch W+2pad, offset pad) for uslice
    bool bulk = (W % 4 == 0) && aligned16(d_in) && (TO_TILES || pad == 0);
    // destination rows: tile rows for slice, ERP rows for uslice
    bool vec = (W % 4 == 0) && aligned16(d_out) && (!TO_TILES || pad == 0);
    int cols4 = (W + 3) / 4;
    int ch = cols4 <= 256 ? 1 : 2;
    int threads = ((cols4 + ch - 1) / ch + 31) / 32 * 32;
    size_t smem = stager_smem(W);
    dim3 grid((unsigned)(planes * sp.strips));
    cudaStream_t s = (cudaStream_t)stream;
#define PCX_RESAMPLE(CHN, MAXT)                                                                               \
    do {                                                                                                      \
        int rc = allow_smem(resample_rows_kernel<CHN, TO_TILES, MAXT>, smem);                                 \
        if (rc) return rc;                                                                                    \
        resample_rows_kernel<CHN, TO_TILES, MAXT><<<grid, threads, smem, s>>>(d_in, d_out, d_src, (const float4 *)d_wt, b, C, \
                                                                               h, W, pad, sp.rows_per_cta, sp.strips, bulk, vec); \
    } while (0)
    if (ch == 1) PCX_RESAMPLE(1, 256);
    else if (threads <= 512) PCX_RESAMPLE(2, 512);
    else PCX_RESAMPLE(2, 1024);
#undef PCX_RESAMPLE
    PCX_LAUNCHED();
    return PCX_OK;
}

template <bool CAUSAL>
int launch_pad(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int pad, const int *wl,
               const int *d_band, const int *d_row, const int *d_col, const float *d_tw, int out_pitch, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_in && d_out && d_band && d_row && d_col && d_tw, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W >= 4, "bad shape N=%d C=%d h=%d W=%d", N, C, h, W);
    PCX_REQUIRE(pad > 0 && pad < 10, "pad %d out of range (pseudo_context_cuda.cu:38)", pad);
    PCX_REQUIRE(C < 1000, "channel count %d >= 1000 (pseudo_context_cuda.cu:38)", C);
    PCX_REQUIRE(out_pitch >= W + 2 * pad, "out_pitch %d < %d", out_pitch, W + 2 * pad);
    PCX_REQUIRE(W <= 8192, "W=%d exceeds the 8192-column limit of the row pipeline", W);
    for (int i = 0; i < npart; i++) PCX_REQUIRE(wl[i] >= 2 * pad && wl[i] <= W, "band %d width %d outside [%d,%d]", i, wl[i], 2 * pad, W);
    i64 planes = (i64)N * npart * C;
    cudaStream_t s = (cudaStream_t)stream;
    // one warp per output row, 8 rows per CTA, a few CTAs per SM (grid-stride over the rows)
    const i64 nrows = planes * (h + 2 * pad);
    i64 want = (nrows + 7) / 8;
    const i64 cap = (i64)pcx_sm_count() * 8;
    const int blocks = (int)(want < cap ? (want < 1 ? 1 : want) : cap);
    pad_direct_kernel<CAUSAL><<<blocks, 256, 0, s>>>(d_in, d_out, b, d_band, d_row, d_col, d_tw, nrows, C, h, W, pad, out_pitch);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // namespace

extern "C" {

int pcx_slice_fwd(const float *d_in, float *d_out, int N, int C, int H, int W, int npart, const int *wl,
                  const int *d_src, const float *d_wt, int pad, void *stream)
{
    PCX_REQUIRE(npart > 0 && H % npart == 0, "height %d is not a multiple of npart %d (math_cuda.cu:225)", H, npart);
    return launch_resample<true>(d_in, d_out, N, C, H / npart, W, npart, wl, d_src, d_wt, pad, stream);
}

int pcx_uslice_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, const int *wl,
                   const int *d_src, const float *d_wt, int pad, void *stream)
{
    return launch_resample<false>(d_in, d_out, N, C, h, W, npart, wl, d_src, d_wt, pad, stream);
}

int pcx_pad_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int pad, const int *wl,
                const int *d_band, const int *d_row, const int *d_col, const float *d_tw, int out_pitch, void *stream)
{
    return launch_pad<false>(d_in, d_out, N, C, h, W, npart, pad, wl, d_band, d_row, d_col, d_tw, out_pitch, stream);
}

int pcx_entropy_pad_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int pad, const int *wl,
                        const int *d_band, const int *d_row, const int *d_col, const float *d_tw, void *stream)
{
    return launch_pad<true>(d_in, d_out, N, C, h, W, npart, pad, wl, d_band, d_row, d_col, d_tw, W + 2 * pad, stream);
}

int pcx_halo_fill(float *d_buf, int N, int C, int h, int W, int npart, int pad, const int *wl, const int *d_band,
                  const int *d_row, const int *d_col, const float *d_tw, int pitch, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_buf && d_band && d_row && d_col && d_tw, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W > 0 && pad > 0 && pitch >= W + 2 * pad, "bad halo_fill geometry");
    i64 planes = (i64)N * npart * C;
    i64 cells = planes * (2 * pad * (W + 2 * pad) + h * 2 * pad);
    int blocks = (int)((cells + 255) / 256 < (i64)pcx_sm_count() * 16 ? (cells + 255) / 256 : (i64)pcx_sm_count() * 16);
    halo_fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_buf, b, d_band, d_row, d_col, d_tw, planes, C, h, W, pad, pitch);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_fill(float *d_data, int N, int C, int Hh, int Ww, int npart, int pad, int trim, const int *wl, float fvalue, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_data && N > 0 && C > 0 && Hh > 0 && Ww > 0 && pad >= 0 && trim >= 0, "bad fill arguments");
    i64 rows = (i64)N * npart * C * Hh;
    i64 want = (rows * 32 + 255) / 256;
    int blocks = (int)(want < (i64)pcx_sm_count() * 16 ? want : (i64)pcx_sm_count() * 16);
    fill_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_data, b, rows, C, Hh, Ww, pad, trim, fvalue);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_dtow(const float *d_in, float *d_out, int N, int C, int H, int W, int stride, int d2w, void *stream)
{
    PCX_REQUIRE(d_in && d_out && N > 0 && C > 0 && H > 0 && W > 0 && stride > 0, "bad dtow arguments");
    if (d2w) PCX_REQUIRE(C % (stride * stride) == 0, "channels %d not divisible by stride^2", C);
    else PCX_REQUIRE(H % stride == 0 && W % stride == 0, "H/W not divisible by stride");
    i64 total = (i64)N * C * H * W;
    const bool al = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    if (stride == 2 && al && total / 4 < (1ll << 31) && (d2w ? W % 2 == 0 : W % 8 == 0)) {
        // d2w: one thread per 4 output columns; w2d: one thread per 8 input columns
        const i64 nvec = d2w ? total / 4 : total / 8;
        i64 want = (nvec + 255) / 256;
        int blocks = (int)(want < (i64)pcx_sm_count() * 16 ? (want < 1 ? 1 : want) : (i64)pcx_sm_count() * 16);
        if (d2w) dtow2_d2w_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, (unsigned)nvec, C, H, W);
        else dtow2_w2d_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, (unsigned)nvec, C, H, W);
        PCX_LAUNCHED();
        return PCX_OK;
    }
    i64 want = (total + 255) / 256;
    int blocks = (int)(want < (i64)pcx_sm_count() * 32 ? want : (i64)pcx_sm_count() * 32);
    dtow_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_in, d_out, total, C, H, W, stride, d2w != 0);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
