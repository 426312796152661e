// pcx_geometry.cu - band geometry, gather tables and wavefront schedules.
//
// Table arithmetic is written with explicit round-to-nearest intrinsics in the operation order the
// reference kernels compile to (nvcc 12.9, -fmad=true, sm_100; verified on SASS), so the tables are
// bit-identical to the reference's without depending on this file's own contraction choices.
#include "pcx_common.cuh"
#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <vector>

// ------------------------------------------------------------------------------------------- library
static thread_local char g_err[512] = "";
std::atomic<long long> g_pcx_launches{0};

void pcx_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void pcx_append_error(const char *fmt, ...)
{
    const size_t n = strlen(g_err);
    if (n + 1 >= sizeof(g_err)) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err + n, sizeof(g_err) - n, fmt, ap);
    va_end(ap);
}

int pcx_sm_count()
{
    static std::atomic<int> cached[64];          // per device: a process may drive several GPUs
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    int n = cached[dev].load(std::memory_order_relaxed);
    if (n == 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev].store(n, std::memory_order_relaxed);
    }
    return n;
}

extern "C" {

int pcx_abi_version(void) { return 2; }
const char *pcx_last_error(void) { return g_err; }
long long pcx_launch_count(void) { return g_pcx_launches.load(); }

int pcx_device_check(int device, int *sm_count, int *cc)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        pcx_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return PCX_ENODEV;
    }
    PCX_REQUIRE(device >= 0 && device < n, "device %d out of range (%d devices)", device, n);
    int major = 0, minor = 0, sms = 0;
    PCX_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    PCX_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
    PCX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = sms;
    if (cc) *cc = major * 10 + minor;
    if (major != 10) {
        pcx_set_error("device %d is sm_%d%d; libpcx is built for sm_100a only", device, major, minor);
        return PCX_ENODEV;
    }
    return PCX_OK;
}

// extension/math_cuda.cu:223-253 (sphere_cal_npart_hw_v3); host, like the reference.
int pcx_band_widths(const float *weight, int npart, int H, int W, int *h_wl)
{
    PCX_REQUIRE(weight && h_wl, "null argument");
    PCX_REQUIRE(npart >= 1 && npart <= PCX_MAX_PART, "npart %d out of range", npart);
    PCX_REQUIRE(H > 0 && W > 0 && H % npart == 0, "height %d must be a positive multiple of npart %d", H, npart);
    const int rows = H / npart;
    float total = 0.f;
    for (int i = 0; i < npart; i++) total += weight[i];
    if (total > 3 * npart) {
        for (int i = 0; i < npart; i++) {
            float scaled = weight[i] / 64 * W;
            h_wl[i] = static_cast<int>(static_cast<double>(scaled) + 0.5);
        }
        return PCX_OK;
    }
    const float pi = static_cast<float>(acos(-1.0));
    const int half = npart / 2;
    auto lat_width = [&](int i, double edge) {
        return static_cast<int>(static_cast<double>(weight[i] * W) * cos((edge / H - 0.5) * static_cast<double>(pi)) + 0.5);
    };
    const int north_end = (npart % 2 == 0) ? half - 1 : half;
    for (int i = 0; i < north_end; i++) h_wl[i] = lat_width(i, rows * (i + 1) - 0.5);
    if (npart % 2 == 0) h_wl[half - 1] = W;
    h_wl[half] = W;
    for (int i = half + 1; i < npart; i++) h_wl[i] = lat_width(i, rows * i + 0.5);
    return PCX_OK;
}
}  // extern "C"

// ------------------------------------------------------------------------------------------- tables
// position of destination sample `dst` (of n_dst) in a row of n_src samples:
//   FUSED: fma(q, n_src, -0.5) + 1e-9   (init_slice_param_kernel / init_uslice_param_kernel)
//   else : q * n_src - 0.5 + 1e-9       (pseudo_context / entropy_context / pseudo_entropy_context kernels)
template <bool FUSED>
__device__ __forceinline__ float resample_position(double dst, int n_dst, int n_src)
{
    double q = __ddiv_rn(__dadd_rn(dst, 0.5), (double)n_dst);
    double v;
    if (FUSED) {
        v = __fma_rn(q, (double)n_src, -0.5);
    } else {
        v = __dmul_rn(q, (double)n_src);
        v = __dadd_rn(v, -0.5);
    }
    v = __dadd_rn(v, 1e-9);
    return __double2float_rn(v);
}

__device__ __forceinline__ float4 catmull_rom(float t)
{
    float t2 = __fmul_rn(t, t);
    float t3 = __fmul_rn(t, t2);
    float a = __fadd_rn(t2, t2);
    a = __fsub_rn(a, t);
    a = __fsub_rn(a, t3);
    float4 w;
    w.x = __fmul_rn(a, 0.5f);
    w.y = __fmul_rn(__fmaf_rn(t3, 3.0f, __fmaf_rn(t2, -5.0f, 2.0f)), 0.5f);
    w.z = __fmul_rn(__fmaf_rn(t3, -3.0f, __fmaf_rn(t2, 4.0f, t)), 0.5f);
    w.w = __fmul_rn(__fsub_rn(t3, t2), 0.5f);
    return w;
}

// to_tiles = true : slice table  (destination = tile column of a band, source = ERP row of W samples)
// to_tiles = false: uslice table (destination = ERP column, source = tile row of wl samples)
__global__ void cubic_table_kernel(Bands bands, int W, int *src, float4 *wt, bool to_tiles)
{
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= bands.npart * W) return;
    int g = idx / W, x = idx % W;
    int wl = bands.wl[g];
    if (to_tiles && x >= wl) {
        src[idx] = 0;
        wt[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    int n_dst = to_tiles ? wl : W;
    int n_src = to_tiles ? W : wl;
    float pos = resample_position<true>((double)x, n_dst, n_src);
    if (pos < 0.f) pos = __fadd_rn(pos, (float)n_src);
    int p = (int)pos;
    float t = __fsub_rn(pos, (float)p);
    src[idx] = p;
    wt[idx] = catmull_rom(t);
}

__global__ void halo_table_kernel(Bands bands, int h, int W, int pad, int mode, int *band, int *row, int *col, float *tw)
{
    i64 idx = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    i64 total = (i64)bands.npart * 2 * pad * W;
    if (idx >= total) return;
    int x = (int)(idx % W);
    int hr = (int)(idx / W);
    int r = hr % pad, s = (hr / pad) % 2, g = hr / pad / 2;
    int Hf = h * bands.npart;
    int ph = (s == 0) ? g * h - pad + r : (g + 1) * h + r;
    bool pole = ph < 0 || ph >= Hf;
    if (pole && mode != 0) {
        if (x == 0) { band[hr] = -1; row[hr] = 0; }
        col[idx] = -1;
        tw[idx] = 0.f;
        return;
    }
    if (pole) ph = ph < 0 ? -ph - 1 : 2 * Hf - ph - 1;
    int pg = ph / h;
    if (x == 0) { band[hr] = pg; row[hr] = ph % h; }
    int wl = bands.wl[g], wsrc = bands.wl[pg];
    if (x >= wl) { col[idx] = 0; tw[idx] = 0.f; return; }
    float pw;
    if (pole) {
        // 180 degree shift across the pole: nw = x + wl/2, wrapped (pseudo_context_cuda.cu:66-69)
        float nw = __double2float_rn(__fma_rn((double)wl, 0.5, (double)x));
        nw = (nw >= (float)wl) ? __fsub_rn(nw, (float)wl) : nw;
        pw = resample_position<false>((double)nw, wl, wsrc);
    } else {
        pw = resample_position<false>((double)x, wl, wsrc);
    }
    if (mode == 0) {
        if (pw < 0.f) pw = __fadd_rn(pw, (float)wsrc);
        int q = (int)pw;
        col[idx] = q;
        tw[idx] = __fsub_rn((float)(q + 1), pw);
    } else if (mode == 1) {
        int q = pw < 0.f ? -1 : (int)pw;
        if (q > x) { col[idx] = -1; tw[idx] = 1.f; }
        else if (q + 1 > x) { col[idx] = q; tw[idx] = 1.f; }
        else { col[idx] = q; tw[idx] = (q == -1) ? 0.f : __fsub_rn((float)(q + 1), pw); }
    } else {
        int q = pw < 0.f ? -1 : (int)pw;
        float t = __fsub_rn((float)(q + 1), pw);
        float qwa = __double2float_rn(__fma_rn(__ddiv_rn((double)(q + 1) + 0.5, (double)wsrc), (double)W, -0.5));
        float qwb = __double2float_rn(__fma_rn(__ddiv_rn((double)x + 0.5, (double)wl), (double)W, -0.5));
        int qidx = (int)qwb;
        if ((double)qwa >= (double)qidx + 0.999) t = 1.f;
        else if (q == -1) t = 0.f;
        col[idx] = q;
        tw[idx] = t;
    }
}

extern "C" {

static int cubic_table(const int *wl, int npart, int W, int *d_src, float *d_wt, void *stream, bool to_tiles)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(W > 0 && d_src && d_wt, "bad table arguments");
    for (int i = 0; i < npart; i++) PCX_REQUIRE(wl[i] >= 4 && wl[i] <= W, "band %d width %d outside [4,%d]", i, wl[i], W);
    int n = npart * W;
    cubic_table_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(b, W, d_src, (float4 *)d_wt, to_tiles);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_slice_table(const int *wl, int npart, int W, int *d_src, float *d_wt, void *stream)
{
    return cubic_table(wl, npart, W, d_src, d_wt, stream, true);
}

int pcx_uslice_table(const int *wl, int npart, int W, int *d_src, float *d_wt, void *stream)
{
    return cubic_table(wl, npart, W, d_src, d_wt, stream, false);
}

int pcx_halo_table(const int *wl, int npart, int h, int W, int pad, int mode, int *d_band, int *d_row, int *d_col,
                   float *d_tw, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(h > 0 && W > 0 && pad > 0 && pad < 10, "bad halo geometry h=%d W=%d pad=%d (pad < 10: pseudo_context_cuda.cu:38)", h, W, pad);
    PCX_REQUIRE(mode >= 0 && mode <= 2, "halo mode %d", mode);
    PCX_REQUIRE(d_band && d_row && d_col && d_tw, "null table pointer");
    i64 n = (i64)npart * 2 * pad * W;
    halo_table_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(b, h, W, pad, mode, d_band, d_row, d_col, d_tw);
    PCX_LAUNCHED();
    return PCX_OK;
}

// entropy_context::reshape_hw, entropy_context_cuda.cu:28-40: valid cells ordered by anti-diagonal.
int pcx_ctx_order(const int *wl, int npart, int h, int W, int *h_order, int *h_start)
{
    PCX_REQUIRE(wl && h_order && h_start && npart >= 1 && npart <= PCX_MAX_PART && h > 0 && W > 0, "bad arguments");
    const int Hf = h * npart;
    // counting pass per plane, then placement: O(cells) instead of the reference's O(planes * rows)
    std::vector<int> cnt(Hf + W, 0);
    for (int i = 0; i < Hf; i++) {
        int w = wl[i / h];
        for (int j = 0; j < w; j++) cnt[i + j + 1]++;
    }
    h_start[0] = 0;
    for (int p = 1; p < Hf + W; p++) h_start[p] = h_start[p - 1] + cnt[p];
    std::vector<int> fillp(h_start, h_start + Hf + W - 1);
    for (int i = 0; i < Hf; i++) {      // rows ascending inside a plane, as the reference emits them
        int w = wl[i / h];
        for (int j = 0; j < w; j++) h_order[fillp[i + j]++] = i * W + j;
    }
    return PCX_OK;
}

// entropy_context_step1/2 + host compaction (entropy_context_cuda.cu:64-103, :187-204).
// record = {kind, a, b, plane}: kind 0 halo (a = table entry, b = halo row), kind 1 right wrap (a = band, b = row*pad+k)
int pcx_ctx_pad_items(const int *wl, int npart, int h, int W, int pad, const int *h_band, const int *h_col,
                      const float *h_tw, int *h_items, int *h_pstart)
{
    PCX_REQUIRE(wl && h_band && h_col && h_tw && h_pstart, "null argument");
    const int Hf = h * npart, nplane = Hf + W + pad - 1;
    std::vector<int> cnt(nplane + 1, 0);
    auto visit = [&](bool emit, std::vector<int> *cursor) {
        for (int g = 0; g < npart; g++) {
            for (int s = 0; s < 2; s++)
                for (int r = 0; r < pad; r++) {
                    int hr = (g * 2 + s) * pad + r;
                    if (h_band[hr] < 0) continue;
                    int ph = (s == 0) ? g * h - pad + r : (g + 1) * h + r;
                    for (int x = 0; x < wl[g]; x++) {
                        i64 e = (i64)hr * W + x;
                        if (h_col[e] < 0 && h_tw[e] >= 1 - 1e-6) continue;
                        int p = ph + x;
                        if (!emit) { cnt[p + 1]++; continue; }
                        int k = (*cursor)[p]++;
                        h_items[k * 4] = 0; h_items[k * 4 + 1] = (int)e; h_items[k * 4 + 2] = hr; h_items[k * 4 + 3] = p;
                    }
                }
            for (int y = 0; y < h + 2 * pad; y++) {
                int ph = g * h + y - pad;
                if (ph < 0 || ph >= Hf) continue;
                for (int k2 = 0; k2 < pad; k2++) {
                    int p = ph + k2 + wl[g];
                    if (!emit) { cnt[p + 1]++; continue; }
                    int k = (*cursor)[p]++;
                    h_items[k * 4] = 1; h_items[k * 4 + 1] = g; h_items[k * 4 + 2] = y * pad + k2; h_items[k * 4 + 3] = p;
                }
            }
        }
    };
    visit(false, nullptr);
    h_pstart[0] = 0;
    for (int p = 1; p <= nplane; p++) h_pstart[p] = h_pstart[p - 1] + cnt[p];
    if (h_items) {
        std::vector<int> cursor(h_pstart, h_pstart + nplane);
        visit(true, &cursor);
    }
    return h_pstart[nplane];
}
}  // extern "C"
