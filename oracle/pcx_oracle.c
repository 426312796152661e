/*
 * pcx_oracle.c - CPU restatement of the reference's codec hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library, and only as the checker / the timed CPU baseline.  The product package
 * (pseudocylindrical_convolution_b200/) never links, imports or calls it.
 *
 * Every function cites the reference file:line (relative to /root/reference) whose arithmetic it
 * restates.  The reference has no CPU implementation of these operators (extension/main.cpp:4-137
 * binds CUDA methods only), so this file is a restatement, not a copy: plain C loops with 64-bit
 * indices and exact integer gather tables (the reference stores integer offsets as float32, which
 * is exact only below 2^24 - SURVEY.md fact 4).  Floating-point expression shapes follow the SASS
 * the reference's kernels compile to with nvcc 12.9 defaults (-fmad=true) for sm_100: where nvcc
 * contracts a*b+c into an FMA this file calls fma()/fmaf() explicitly, and it must be built with
 * -ffp-contract=off so the C compiler adds no contraction of its own.
 *
 * Parity pinning: the reference ships no golden vectors for this path.  The
 * oracle is pinned against outputs of the reference extension itself, built unmodified for sm_100
 * (oracle/build_ref.py -> oracle/_ref/PCONV_ref.so) and run on the B200 box; the vectors it produced
 * are committed under tests/golden/ with the script that made them (tools/make_golden.py + tests/golden_cases.py).
 * Known, documented non-bit-exact spots versus the GPU: libm erff/expf (CUDA libdevice differs from
 * glibc by <= 1-2 ulp) - CDF entries may differ by +-1 count, quantiser step tables by <= 2 ulp.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))
typedef int64_t i64;


/* ------------------------------------------------------------------------------------------------
 * Geometry.  extension/math_cuda.cu:223-253 (sphere_cal_npart_hw_v3); v2 (:177-221) yields the same
 * widths.  weight[] is PCONV_operator/base.py:13-35 set_weight().
 * ---------------------------------------------------------------------------------------------- */
API int orc_band_widths(const float *weight, int npart, int H, int W, int *wl)
{
    if (H % npart != 0) return -1;                       /* math_cuda.cu:225 assert */
    int hpp = H / npart;
    float total = 0;
    for (int i = 0; i < npart; i++) total += weight[i];
    if (total > 3 * npart) {                             /* :230-235 "opt" profile, units of 1/64 */
        for (int i = 0; i < npart; i++) {
            float f = weight[i] / 64 * W;                /* float product */
            wl[i] = (int)((double)f + 0.5);              /* +0.5 promotes to double */
        }
        return 0;
    }
    float pi = (float)acos(-1.0);                        /* :237 float pi */
    int half = npart / 2;
    if (npart % 2 == 0) {
        for (int i = 0; i < half - 1; i++)
            wl[i] = (int)((double)(weight[i] * W) * cos(((hpp * (i + 1) - 0.5) / H - 0.5) * (double)pi) + 0.5);
        wl[half - 1] = W;
        wl[half] = W;
        for (int i = half + 1; i < npart; i++)
            wl[i] = (int)((double)(weight[i] * W) * cos(((hpp * i + 0.5) / H - 0.5) * (double)pi) + 0.5);
    } else {
        for (int i = 0; i < half; i++)
            wl[i] = (int)((double)(weight[i] * W) * cos(((hpp * (i + 1) - 0.5) / H - 0.5) * (double)pi) + 0.5);
        wl[half] = W;
        for (int i = half + 1; i < npart; i++)
            wl[i] = (int)((double)(weight[i] * W) * cos(((hpp * i + 0.5) / H - 0.5) * (double)pi) + 0.5);
    }
    return 0;
}

/* Catmull-Rom weights exactly as init_slice_param_kernel<float> compiles (sphere_slice_cuda.cu:21-29):
 * t2=t*t, t3=t*t2 (FMUL); w0=((t2+t2)-t-t3)*.5; w1=fma(t3,3,fma(t2,-5,2))*.5;
 * w2=fma(t3,-3,fma(t2,4,t))*.5; w3=(t3-t2)*.5. */
static void cubic_weights(float t, float *w)
{
    float t2 = t * t;
    float t3 = t * t2;
    float a = t2 + t2;
    a = a - t;
    a = a - t3;
    w[0] = a * 0.5f;
    w[1] = fmaf(t3, 3.0f, fmaf(t2, -5.0f, 2.0f)) * 0.5f;
    w[2] = fmaf(t3, -3.0f, fmaf(t2, 4.0f, t)) * 0.5f;
    w[3] = (t3 - t2) * 0.5f;
}

/* (dst + 0.5) / n_dst * n_src - 0.5 + 1e-9 in double, rounded once to float.  SASS of
 * init_slice_param_kernel / init_uslice_param_kernel (sphere_slice_cuda.cu:19, sphere_uslice_cuda.cu:18):
 * IEEE double division, DFMA(q, n_src, -0.5), DADD 1e-9, F2F.F32.F64. */
static float resample_pos(double dst, int n_dst, int n_src)
{
    double q = (dst + 0.5) / (double)n_dst;
    double v = fma(q, (double)n_src, -0.5);
    v = v + 1e-9;
    return (float)v;
}

/* Same expression in the halo-table kernels (pseudo_context_cuda.cu:70,73; entropy_context_cuda.cu:125,133;
 * pseudo_entropy_context_cuda.cu:70,78,131,139) compiles WITHOUT contraction: DMUL, DADD -0.5, DADD 1e-9. */
static float resample_pos_halo(double dst, int n_dst, int n_src)
{
    double q = (dst + 0.5) / (double)n_dst;
    double v = q * (double)n_src;
    v = v - 0.5;
    v = v + 1e-9;
    return (float)v;
}

/* sphere_slice_cuda.cu:13-32.  src[g*W+x] = integer tap base, wt[(g*W+x)*4..] = 4 weights. */
API void orc_slice_table(const int *wl, int npart, int W, int *src, float *wt)
{
    for (int g = 0; g < npart; g++)
        for (int x = 0; x < W; x++) {
            i64 k = (i64)g * W + x;
            src[k] = 0;
            wt[k * 4] = wt[k * 4 + 1] = wt[k * 4 + 2] = wt[k * 4 + 3] = 0.f;
            if (x >= wl[g]) continue;
            float nidx = resample_pos((double)x, wl[g], W);
            if (nidx < 0) nidx = nidx + (float)W;
            int p = (int)nidx;
            float t = nidx - (float)p;
            src[k] = p;
            cubic_weights(t, wt + k * 4);
        }
}

/* sphere_uslice_cuda.cu:13-30. */
API void orc_uslice_table(const int *wl, int npart, int W, int *src, float *wt)
{
    for (int g = 0; g < npart; g++)
        for (int x = 0; x < W; x++) {
            i64 k = (i64)g * W + x;
            float nidx = resample_pos((double)x, W, wl[g]);
            if (nidx < 0) nidx = nidx + (float)wl[g];
            int p = (int)nidx;
            float t = nidx - (float)p;
            src[k] = p;
            cubic_weights(t, wt + k * 4);
        }
}

/* 4-tap accumulate.  nvcc chooses which product is the plain FMUL and chooses differently in the two kernels
 * (verified on the reference's SASS, nvcc 12.9 sm_100):
 *   slice  (sphere_slice_cuda.cu:109-113): FMUL p2*i2, then FFMA p1*i1, p3*i3, p4*i4
 *   uslice (sphere_uslice_cuda.cu:91-96) : FMUL p1*i1, then FFMA p2*i2, p3*i3, p4*i4 */
static inline float tap4_slice(const float *w, float a, float b, float c, float d)
{
    float r = w[1] * b;
    r = fmaf(w[0], a, r);
    r = fmaf(w[2], c, r);
    r = fmaf(w[3], d, r);
    return r;
}

static inline float tap4_uslice(const float *w, float a, float b, float c, float d)
{
    float r = w[0] * a;
    r = fmaf(w[1], b, r);
    r = fmaf(w[2], c, r);
    r = fmaf(w[3], d, r);
    return r;
}

/* sphere_slice_cuda.cu:87-116.  in (N,C,H,W) -> out (N*npart, C, h+2pad, W+2pad); only the interior
 * is written (the reference leaves the border of its at::empty output untouched); pad=0 in the codec. */
API void orc_slice(const float *in, float *out, int N, int C, int H, int W, int npart, const int *wl,
                   const int *src, const float *wt, int pad)
{
    int h = H / npart;
    i64 oh = h + 2 * pad, ow = W + 2 * pad;
    for (int n = 0; n < N; n++)
        for (int g = 0; g < npart; g++)
            for (int c = 0; c < C; c++)
                for (int y = 0; y < h; y++) {
                    const float *row = in + (((i64)n * C + c) * H + (i64)g * h + y) * W;
                    float *o = out + ((((i64)n * npart + g) * C + c) * oh + y + pad) * ow + pad;
                    for (int x = 0; x < W; x++) {
                        if (x >= wl[g]) { o[x] = 0.f; continue; }
                        i64 k = (i64)g * W + x;
                        int p = src[k];
                        o[x] = tap4_slice(wt + k * 4, row[(p - 1 + W) % W], row[p], row[(p + 1) % W], row[(p + 2) % W]);
                    }
                }
}

/* sphere_uslice_cuda.cu:73-99.  in (N*npart, C, h+2pad, W+2pad) -> out (N, C, h*npart, W). */
API void orc_uslice(const float *in, float *out, int N, int C, int h, int W, int npart, const int *wl,
                    const int *src, const float *wt, int pad)
{
    i64 ih = h + 2 * pad, iw = W + 2 * pad;
    for (int n = 0; n < N; n++)
        for (int c = 0; c < C; c++)
            for (int g = 0; g < npart; g++)
                for (int y = 0; y < h; y++) {
                    const float *row = in + ((((i64)n * npart + g) * C + c) * ih + y + pad) * iw + pad;
                    float *o = out + (((i64)n * C + c) * ((i64)h * npart) + (i64)g * h + y) * W;
                    int w = wl[g];
                    for (int x = 0; x < W; x++) {
                        i64 k = (i64)g * W + x;
                        int p = src[k];
                        o[x] = tap4_uslice(wt + k * 4, row[(p - 1 + w) % w], row[p], row[(p + 1) % w], row[(p + 2) % w]);
                    }
                }
}

/* ------------------------------------------------------------------------------------------------
 * Halo tables.  One entry per (band g, side s in {0=top,1=bottom}, halo row r < pad, column x < wl[g]).
 *   mode 0: transform pad   - pseudo_context_cuda.cu:51-104 (poles mirror + 180 degree shift)
 *   mode 1: causal pad      - entropy_context_cuda.cu:105-165 == pseudo_entropy_context_cuda.cu:111-170 (v1)
 *   mode 2: causal pad v0   - pseudo_entropy_context_cuda.cu:50-109
 * Outputs: band[(g*2+s)*pad+r] = source band or -1 (pole, modes 1/2); row[...] = source row inside
 * that band; col[e] = left source column (may be -1 in causal modes); tw[e] = weight of the left
 * sample, e = ((g*2+s)*pad+r)*W + x.
 * ---------------------------------------------------------------------------------------------- */
API void orc_halo_table(const int *wl, int npart, int h, int W, int pad, int mode,
                        int *band, int *row, int *col, float *tw)
{
    int Hf = h * npart;
    for (int g = 0; g < npart; g++)
        for (int s = 0; s < 2; s++)
            for (int r = 0; r < pad; r++) {
                int hr = (g * 2 + s) * pad + r;
                int ph = (s == 0) ? g * h - pad + r : (g + 1) * h + r;
                int pole = (ph < 0) || (ph >= Hf);
                int pg;
                if (pole && mode != 0) {
                    band[hr] = -1;
                    row[hr] = 0;
                    for (int x = 0; x < W; x++) { col[(i64)hr * W + x] = -1; tw[(i64)hr * W + x] = 0.f; }
                    continue;
                }
                if (pole) ph = (ph < 0) ? -ph - 1 : 2 * Hf - ph - 1;
                pg = ph / h;
                band[hr] = pg;
                row[hr] = ph % h;
                for (int x = 0; x < W; x++) {
                    i64 e = (i64)hr * W + x;
                    col[e] = 0; tw[e] = 0.f;
                    if (x >= wl[g]) continue;
                    float pw;
                    if (pole) {
                        /* nw = tw + wl/2. (double, exact) stored to float; wrap; pseudo_context_cuda.cu:66-70 */
                        float nw = (float)fma((double)wl[g], 0.5, (double)x);
                        nw = (nw >= (float)wl[g]) ? nw - (float)wl[g] : nw;
                        pw = resample_pos_halo((double)nw, wl[g], wl[pg]);
                    } else {
                        pw = resample_pos_halo((double)x, wl[g], wl[pg]);
                    }
                    if (mode == 0) {
                        if (pw < 0) pw = pw + (float)wl[pg];                /* :91 */
                        int q = (int)pw;
                        col[e] = q;
                        tw[e] = (float)(q + 1) - pw;                        /* :95 */
                    } else if (mode == 1) {
                        int q = pw < 0 ? -1 : (int)pw;                      /* entropy_context_cuda.cu:146 */
                        if (q > x) { col[e] = -1; tw[e] = 1.f; }
                        else if (q + 1 > x) { col[e] = q; tw[e] = 1.f; }
                        else { col[e] = q; tw[e] = (q == -1) ? 0.f : (float)(q + 1) - pw; }
                    } else {
                        int q = pw < 0 ? -1 : (int)pw;                      /* pseudo_entropy_context_cuda.cu:86-100 */
                        col[e] = q;
                        float t = (float)(q + 1) - pw;
                        float qwa = (float)fma((q + 1 + 0.5) / (double)wl[pg], (double)W, -0.5);   /* DFMA in SASS */
                        float qwb = (float)fma((x + 0.5) / (double)wl[g], (double)W, -0.5);
                        int qidx = (int)qwb;
                        if ((double)qwa >= qidx + 0.999) t = 1.f;
                        else if (q == -1) t = 0.f;
                        tw[e] = t;
                    }
                }
            }
}

/* 2-tap halo interpolation as pseudo_pad_forward_kernel<float> compiles (pseudo_pad.cu:77):
 * FADD r=1-t; FMUL r=b*r; FFMA out=a*t+r. */
static inline float lerp2(float a, float b, float t)
{
    float r = 1.0f - t;
    r = b * r;
    return fmaf(a, t, r);
}

/* pseudo_pad.cu:39-96 (three kernels) restated as one pass per output row.
 * in (N*npart, C, h, W) -> out (N*npart, C, h+2p, W+2p).  Table from orc_halo_table(mode 0). */
API void orc_pad(const float *in, float *out, int N, int C, int h, int W, int npart, int pad,
                 const int *wl, const int *band, const int *row, const int *col, const float *tw)
{
    i64 oh = h + 2 * pad, ow = W + 2 * pad;
    for (int n = 0; n < N; n++)
        for (int g = 0; g < npart; g++)
            for (int c = 0; c < C; c++) {
                float *op = out + (((i64)n * npart + g) * C + c) * oh * ow;
                int w = wl[g];
                for (int y = 0; y < oh; y++) {
                    float *o = op + (i64)y * ow;
                    for (int x = 0; x < ow; x++) o[x] = 0.f;
                    if (y >= pad && y < pad + h) {
                        const float *src = in + ((((i64)n * npart + g) * C + c) * h + (y - pad)) * W;
                        for (int x = 0; x < w; x++) o[pad + x] = src[x];
                    } else {
                        int s = y < pad ? 0 : 1;
                        int r = y < pad ? y : y - pad - h;
                        int hr = (g * 2 + s) * pad + r;
                        int pg = band[hr];
                        const float *src = in + ((((i64)n * npart + pg) * C + c) * h + row[hr]) * W;
                        for (int x = 0; x < w; x++) {
                            i64 e = (i64)hr * W + x;
                            int q = col[e];
                            o[pad + x] = lerp2(src[q], src[(q + 1) % wl[pg]], tw[e]);
                        }
                    }
                    /* longitude wrap, every row (pseudo_pad.cu:82-96) */
                    for (int k = 0; k < pad; k++) {
                        o[k] = o[pad + w - pad + k];
                        o[pad + w + k] = o[pad + k];
                    }
                }
            }
}

/* Full-tensor causal pad (training-path form of the context pad): pseudo_entropy_pad_cuda.cu:39-105.
 * Table from orc_halo_table(mode 1 or 2).  Left pad columns = 0, right pad = wrap, pole rows = 0. */
API void orc_entropy_pad(const float *in, float *out, int N, int C, int h, int W, int npart, int pad,
                         const int *wl, const int *band, const int *row, const int *col, const float *tw)
{
    i64 oh = h + 2 * pad, ow = W + 2 * pad;
    for (int n = 0; n < N; n++)
        for (int g = 0; g < npart; g++)
            for (int c = 0; c < C; c++) {
                float *op = out + (((i64)n * npart + g) * C + c) * oh * ow;
                int w = wl[g];
                for (int y = 0; y < oh; y++) {
                    float *o = op + (i64)y * ow;
                    for (int x = 0; x < ow; x++) o[x] = 0.f;
                    if (y >= pad && y < pad + h) {
                        const float *src = in + ((((i64)n * npart + g) * C + c) * h + (y - pad)) * W;
                        for (int x = 0; x < w; x++) o[pad + x] = src[x];
                    } else {
                        int s = y < pad ? 0 : 1;
                        int r = y < pad ? y : y - pad - h;
                        int hr = (g * 2 + s) * pad + r;
                        int pg = band[hr];
                        if (pg >= 0) {
                            const float *src = in + ((((i64)n * npart + pg) * C + c) * h + row[hr]) * W;
                            for (int x = 0; x < w; x++) {
                                i64 e = (i64)hr * W + x;
                                int q = col[e];
                                float a = (q == -1) ? 0.f : src[q];
                                o[pad + x] = lerp2(a, src[(q + 1) % wl[pg]], tw[e]);
                            }
                        }
                    }
                    for (int k = 0; k < pad; k++) {
                        o[k] = 0.f;
                        o[pad + w + k] = o[pad + k];
                    }
                }
            }
}

/* pseudo_fill_cuda.cu:28-43.  In place on (N*npart, C, Hh, Ww) where Hh/Ww include `pad`. */
API void orc_fill(float *data, int N, int C, int Hh, int Ww, int npart, int pad, int trim,
                  const int *wl, float fvalue)
{
    for (i64 nn = 0; nn < (i64)N * npart; nn++) {
        int g = (int)(nn % npart);
        for (int c = 0; c < C; c++)
            for (int y = 0; y < Hh; y++) {
                float *o = data + ((nn * C + c) * Hh + y) * (i64)Ww;
                int rowout = (y < pad - trim) || (y >= Hh - pad + trim);
                for (int x = 0; x < Ww; x++)
                    if (rowout || x < pad - trim || x >= pad + wl[g] + trim) o[x] = fvalue;
            }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Quantiser.  pseudo_quant_cuda.cu:37-94, pseudo_dquant_cuda.cu:24-47.
 * ---------------------------------------------------------------------------------------------- */
API void orc_quant_steps(const float *theta, float *w, int C, int L)     /* :37-45 */
{
    for (i64 i = 0; i < (i64)C * L; i++) w[i] = (i % L == 0) ? theta[i] : expf(theta[i]);
}

API void orc_dquant_centres(const float *theta, float *cen, int C, int L) /* pseudo_dquant_cuda.cu:24-31 */
{
    for (int c = 0; c < C; c++) {
        cen[(i64)c * L] = theta[(i64)c * L];
        for (int j = 1; j < L; j++) cen[(i64)c * L + j] = cen[(i64)c * L + j - 1] + expf(theta[(i64)c * L + j]);
    }
}

/* w = expanded step table (orc_quant_steps or the table the GPU computed).  count accumulates -1 per
 * symbol (the reference's histogram side effect, float atomics, :65/:83). */
API void orc_quant(const float *x, float *val, float *sym, float *count, const float *w,
                   int N, int C, int h, int W, int npart, int L, const int *wl)
{
    for (i64 nn = 0; nn < (i64)N * npart; nn++) {
        int g = (int)(nn % npart);
        for (int c = 0; c < C; c++)
            for (int y = 0; y < h; y++)
                for (int xx = 0; xx < W; xx++) {
                    i64 i = ((nn * C + c) * h + y) * (i64)W + xx;
                    if (xx >= wl[g]) { val[i] = 0.f; sym[i] = 0.f; continue; }
                    const float *wc = w + (i64)c * L;
                    float tmp = x[i] - wc[0];
                    if (tmp < 0) { sym[i] = 0.f; val[i] = wc[0]; if (count) count[(i64)c * L] -= 1.f; continue; }
                    int j = 1;
                    for (; j < L; j++) { tmp -= wc[j]; if (tmp < 0) break; }
                    if (j == L) j--;
                    if (tmp + tmp + wc[j] < 0) { tmp = tmp + wc[j]; j--; }
                    val[i] = x[i] - tmp;
                    sym[i] = (float)j;
                    if (count) count[(i64)c * L + j] -= 1.f;
                }
    }
}

API void orc_dquant(const float *sym, float *out, const float *cen, int N, int C, int h, int W,
                    int npart, int L, const int *wl)
{
    for (i64 nn = 0; nn < (i64)N * npart; nn++) {
        int g = (int)(nn % npart);
        for (int c = 0; c < C; c++)
            for (int y = 0; y < h; y++)
                for (int xx = 0; xx < W; xx++) {
                    i64 i = ((nn * C + c) * h + y) * (i64)W + xx;
                    if (xx >= wl[g]) { out[i] = 0.f; continue; }
                    int idx = (int)((double)sym[i] + 0.00001);
                    out[i] = cen[(i64)c * L + idx];
                }
    }
}

/* dtow_cuda.cu:38-75: depth->space (d2w) / space->depth, patch size s. */
API void orc_dtow(const float *in, float *out, int N, int C, int H, int W, int s, int d2w)
{
    int s2 = s * s;
    if (d2w) {
        int Co = C / s2; i64 Ho = (i64)H * s, Wo = (i64)W * s;
        for (int n = 0; n < N; n++) for (int c = 0; c < C; c++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            int pc = c / s2, rc = c % s2;
            i64 py = (i64)y * s + rc / s, px = (i64)x * s + rc % s;
            out[(((i64)n * Co + pc) * Ho + py) * Wo + px] = in[(((i64)n * C + c) * H + y) * W + x];
        }
    } else {
        int Co = C * s2; i64 Ho = H / s, Wo = W / s;
        for (int n = 0; n < N; n++) for (int c = 0; c < C; c++) for (int y = 0; y < H; y++) for (int x = 0; x < W; x++) {
            int pc = c * s2 + (y % s) * s + x % s;
            out[(((i64)n * Co + pc) * Ho + y / s) * Wo + x / s] = in[(((i64)n * C + c) * H + y) * W + x];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Wavefront context model.  entropy_context_cuda.cu:13-45 (order), :64-103 + :187-204 (per-plane
 * halo / right-wrap work lists).  Work items are emitted in a deterministic order (the reference's
 * order inside a plane depends on atomics and does not affect results).
 * ---------------------------------------------------------------------------------------------- */
API void orc_ctx_order(const int *wl, int npart, int h, int W, int *order, int *start)
{
    int Hf = h * npart, k = 0;
    for (int p = 0; p < Hf + W - 1; p++) {
        start[p] = k;
        for (int i = 0; i < Hf; i++) {
            int j = p - i;
            if (j < 0 || j >= wl[i / h]) continue;
            order[k++] = i * W + j;
        }
    }
    start[Hf + W - 1] = k;
}

/* Items: int4 {kind, a, b, plane}. kind 0 = halo (a = table entry e, b = halo-row hr),
 * kind 1 = right wrap (a = band g, b = padded row y * pad + k).  Returns item count; pstart has
 * Hf + W + pad entries (prefix over planes 0 .. Hf+W+pad-2). */
API int orc_ctx_pad_items(const int *wl, int npart, int h, int W, int pad,
                          const int *band, const int *col, const float *tw, int *items, int *pstart)
{
    int Hf = h * npart, nplane = Hf + W + pad - 1, n = 0;
    for (int p = 0; p < nplane; p++) {
        pstart[p] = n;
        for (int g = 0; g < npart; g++) {
            for (int s = 0; s < 2; s++)
                for (int r = 0; r < pad; r++) {
                    int hr = (g * 2 + s) * pad + r;
                    if (band[hr] < 0) continue;
                    int ph = (s == 0) ? g * h - pad + r : (g + 1) * h + r;
                    int x = p - ph;
                    if (x < 0 || x >= wl[g]) continue;
                    i64 e = (i64)hr * W + x;
                    if (col[e] < 0 && tw[e] >= 1 - 1e-6) continue;       /* entropy_context_cuda.cu:74 */
                    if (items) { items[n * 4] = 0; items[n * 4 + 1] = (int)e; items[n * 4 + 2] = hr; items[n * 4 + 3] = p; }
                    n++;
                }
            for (int y = 0; y < h + 2 * pad; y++) {
                int ph = g * h + y - pad;
                if (ph < 0 || ph >= Hf) continue;
                int k = p - ph - wl[g];
                if (k < 0 || k >= pad) continue;
                if (items) { items[n * 4] = 1; items[n * 4 + 1] = g; items[n * 4 + 2] = y * pad + k; items[n * 4 + 3] = p; }
                n++;
            }
        }
    }
    pstart[nplane] = n;
    return n;
}

/* entropy_ctx_pad_run2_cuda.cu:33-65, :86-117.  buf (nrep*npart, G*cpn, h+2p, W+2p) in place;
 * `psum` is the step AFTER the input-layer lag has been applied by the caller (the op subtracts 1
 * when input_ is set, :93-95). */
API void orc_ctx_pad_step(float *buf, int nrep, int npart, int G, int cpn, int h, int W, int pad, int psum,
                          const int *wl, const int *band, const int *row, const int *col, const float *tw,
                          const int *items, const int *pstart)
{
    int Hf = h * npart;
    int mod = Hf + W + pad + G - 2;
    if (psum < 0 || psum >= mod) return;
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W + pad - 2 ? psum + 1 : Hf + W + pad - 1;
    i64 oh = h + 2 * pad, ow = W + 2 * pad, C = (i64)G * cpn;
    for (int it = pstart[st]; it < pstart[en]; it++) {
        const int *I = items + (i64)it * 4;
        int grp = psum - I[3];
        for (int n = 0; n < nrep; n++)
            for (int cc = 0; cc < cpn; cc++) {
                i64 c = (i64)grp * cpn + cc;
                if (I[0] == 0) {
                    int e = I[1], hr = I[2];
                    int g = hr / (2 * pad), s = (hr / pad) % 2, r = hr % pad, x = e % W;
                    int y = s == 0 ? r : pad + h + r;
                    int pg = band[hr];
                    float *dst = buf + ((((i64)n * npart + g) * C + c) * oh + y) * ow + pad + x;
                    const float *src = buf + ((((i64)n * npart + pg) * C + c) * oh + pad + row[hr]) * ow + pad;
                    int q = col[e];
                    float a = (q == -1) ? 0.f : src[q];
                    *dst = lerp2(a, src[(q + 1) % wl[pg]], tw[e]);
                } else {
                    int g = I[1], y = I[2] / pad, k = I[2] % pad;
                    float *rowp = buf + ((((i64)n * npart + g) * C + c) * oh + y) * ow;
                    rowp[pad + wl[g] + k] = rowp[pad + k];
                }
            }
    }
}

/* One output scalar of the masked grouped 5x5 convolution with the reference's reduction tree:
 * entropy_conv_cuda_v2.cu:326-379 (act_batch) / :237-290 (batch).  128 virtual threads; thread i <
 * 25*gi owns tap (kw=i%5, kh=(i/5)%5, m=i/25) and accumulates FFMA over channel groups in ascending
 * order; then [t]+=[t+64], [t]+=[t+32], shuffle-down 16,8,4,2,1. */
static float ctx_conv_scalar(const float *in_cell, const float *wgt, int gi, int nallow_base, int G,
                             i64 in_cstride, i64 in_rstride, int hp, int tw_, int psum, int constrain)
{
    float v[128];
    for (int i = 0; i < 128; i++) v[i] = 0.f;
    int nth = 25 * gi;
    for (int i = 0; i < nth && i < 128; i++) {
        int kw = i % 5, kh = (i / 5) % 5, m = i / 25;
        int qh = hp - 2 + kh, pw = tw_ - 2 + kw;
        int nch = (constrain == 5 ? (psum - qh - pw) : (psum - qh - pw + 1)) * gi;
        if (nch > G * gi) nch = G * gi;
        float acc = 0.f;
        if (nch > 0) {
            const float *ip = in_cell + (i64)(kh - 2) * in_rstride + (kw - 2);
            const float *wp = wgt + kh * 5 + kw;
            for (int ti = m; ti < nch; ti += gi) acc = fmaf(ip[(i64)ti * in_cstride], wp[(i64)ti * 25], acc);
        }
        v[i] = acc;
    }
    (void)nallow_base;
    for (int t = 0; t < 64; t++) v[t] = v[t] + v[t + 64];
    for (int t = 0; t < 32; t++) v[t] = v[t] + v[t + 32];
    for (int off = 16; off > 0; off >>= 1)
        for (int t = 0; t < off; t++) v[t] = v[t] + v[t + off];
    return v[0];
}

/* in  (nb*nimg*npart, G*gi, h+2pi, W+2pi), weight (nb, G*go, G*gi, 5, 5), bias/act (nb, G*go),
 * out (nb*nimg*npart, G*go, h+2po, W+2po).  act == NULL -> no PReLU. */
/* The output scalars of one step are independent: tasks [task_lo, task_hi) of the nb*nimg*cells (net image, cell) pairs, so
 * that the CPU baseline (bench.py --impl reference) can spread a step over the host cores (oracle.py splits the range over a
 * thread pool; every scalar keeps its own reduction order).  task_hi < 0 = all. */
API void orc_ctx_conv_step_range(const float *in, const float *weight, const float *bias, const float *act,
                                 float *out, int nb, int nimg, int npart, int G, int gi, int go, int h, int W,
                                 int pad_in, int pad_out, int constrain, int psum, const int *order, const int *start,
                                 i64 task_lo, i64 task_hi)
{
    int Hf = h * npart;
    int mod = Hf + W + G - 2;
    if (psum >= mod) return;
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W - 2 ? psum + 1 : Hf + W - 1;
    i64 ih = h + 2 * pad_in, iw = W + 2 * pad_in, oh = h + 2 * pad_out, ow = W + 2 * pad_out;
    i64 Ci = (i64)G * gi, Co = (i64)G * go;
    const int ncell = start[en] - start[st];
    const i64 ntask = (i64)nb * nimg * ncell;
    if (task_hi < 0 || task_hi > ntask) task_hi = ntask;
    for (i64 task = task_lo < 0 ? 0 : task_lo; task < task_hi; task++) {
        int pn = (int)(task / ncell), k = start[st] + (int)(task % ncell);
        int b = pn / nimg;
        int hw = order[k], tw_ = hw % W, hp = hw / W, g = hp / h, th = hp % h;
        int tc = psum - tw_ - hp;
        i64 qn = (i64)pn * npart + g;
        const float *in_cell = in + (qn * Ci * ih + th + pad_in) * iw + tw_ + pad_in;
        for (int og = 0; og < go; og++) {
            int pout = tc * go + og;
            const float *wgt = weight + ((i64)b * Co + pout) * Ci * 25;
            float s = ctx_conv_scalar(in_cell, wgt, gi, 0, G, ih * iw, iw, hp, tw_, psum, constrain);
            s = s + bias[(i64)b * Co + pout];
            if (act && s < 0) s = s * act[(i64)b * Co + pout];
            out[((qn * Co + pout) * oh + th + pad_out) * ow + tw_ + pad_out] = s;
        }
    }
}

API void orc_ctx_conv_step(const float *in, const float *weight, const float *bias, const float *act,
                           float *out, int nb, int nimg, int npart, int G, int gi, int go, int h, int W,
                           int pad_in, int pad_out, int constrain, int psum, const int *order, const int *start)
{
    orc_ctx_conv_step_range(in, weight, bias, act, out, nb, nimg, npart, G, gi, go, h, W, pad_in, pad_out, constrain, psum, order,
                            start, 0, -1);
}

/* entropy_add_cuda.cu:25-44: y += x at the wavefront cells (both (nrep*npart, G*cpg, h+2p, W+2p)). */
API void orc_ctx_add_step(float *y, const float *x, int nrep, int npart, int G, int cpg, int h, int W,
                          int pad, int psum, const int *order, const int *start)
{
    int Hf = h * npart;
    if (psum > Hf + W + G - 2) return;
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W - 2 ? psum + 1 : Hf + W - 1;
    i64 oh = h + 2 * pad, ow = W + 2 * pad, C = (i64)G * cpg;
    for (int pn = 0; pn < nrep; pn++)
        for (int k = start[st]; k < start[en]; k++) {
            int hw = order[k], tw_ = hw % W, hp = hw / W, g = hp / h, th = hp % h;
            int tc = psum - tw_ - hp;
            for (int og = 0; og < cpg; og++) {
                i64 i = ((((i64)pn * npart + g) * C + (i64)tc * cpg + og) * oh + th + pad) * ow + tw_ + pad;
                y[i] = y[i] + x[i];
            }
        }
}

/* d_input_cuda_v2.cu:32-52, :55-86.  `psum` is the op's raw counter; the scatter uses psum-1.
 * sym: compact list (nimg, len) of the previous step's symbols.  out (rep*nimg*npart, G, h+2p, W+2p). */
API void orc_dinput_step(const float *sym, float *out, int nimg, int npart, int G, int h, int W, int pad,
                         float bias, int rep, int psum, const int *order, const int *start)
{
    int Hf = h * npart;
    i64 oh = h + 2 * pad, ow = W + 2 * pad;
    i64 rep_stride = (i64)nimg * npart * G * oh * ow;
    if (psum == 0) { memset(out, 0, sizeof(float) * rep * rep_stride); return; }
    if (psum > Hf + W + G - 2) return;
    psum -= 1;
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W - 2 ? psum + 1 : Hf + W - 1;
    int len = start[en] - start[st];
    for (int n = 0; n < nimg; n++)
        for (int k = 0; k < len; k++) {
            int hw = order[start[st] + k], tw_ = hw % W, hp = hw / W, g = hp / h, th = hp % h;
            int tc = psum - tw_ - hp;
            float v = sym[(i64)n * len + k] + bias;
            i64 i = ((((i64)n * npart + g) * G + tc) * oh + th + pad) * ow + tw_ + pad;
            for (int j = 0; j < rep; j++) out[i + j * rep_stride] = v;
        }
}

/* d_extract_cuda_v2.cu:34-52 (label / plain gather), :110-132 (batch gather of the 3 nets).
 * in (nrep*npart, G*cpn, h, W) unpadded.  Returns the number of rows written.
 *   batch == 0: out[(n*len + k)*cpn + ci], nrep images.
 *   batch == 1: nrep = 3*nimg; out[net*stride + ((img*len + k)*cpn + ci)], stride = cpn*Hf*W*nimg. */
API int orc_dextract_step(const float *in, float *out, int nrep, int npart, int G, int cpn, int h, int W,
                          int psum, int batch, const int *order, const int *start)
{
    int Hf = h * npart;
    if (psum >= Hf + W + G - 2) return 0;
    int st = psum - G + 1 < 0 ? 0 : psum - G + 1;
    int en = psum < Hf + W - 2 ? psum + 1 : Hf + W - 1;
    int len = start[en] - start[st];
    i64 C = (i64)G * cpn;
    int nimg = batch ? nrep / 3 : nrep;
    i64 stride = (i64)cpn * Hf * W * nimg;
    for (int n = 0; n < nrep; n++)
        for (int k = 0; k < len; k++) {
            int hw = order[start[st] + k], tw_ = hw % W, hp = hw / W, g = hp / h, th = hp % h;
            int tc = psum - tw_ - hp;
            for (int ci = 0; ci < cpn; ci++) {
                float v = in[((((i64)n * npart + g) * C + (i64)tc * cpn + ci) * h + th) * W + tw_];
                if (batch) out[(n / nimg) * stride + (((i64)(n % nimg) * len + k) * cpn + ci)] = v;
                else out[((i64)n * len + k) * cpn + ci] = v;
            }
        }
    return batch ? nimg * len : nrep * len;
}

/* ------------------------------------------------------------------------------------------------
 * GMM integer CDF tables.  entropy_gmm_table_cuda.cu:29-47 (softmax), :50-56 (delta), :136-153 (CDF),
 * :83-105 (monotonic fix-up).  logit/delta/mean: (n, ng) rows.  cdf: (n, nstep+1) int32.
 * libm note: the GPU uses libdevice expf/erff (erff's large-|x| branch goes through MUFU.EX2); glibc
 * differs by <= 1-2 ulp, so cdf entries may differ from the GPU by +-1 on rare rows.
 * ---------------------------------------------------------------------------------------------- */
API void orc_gmm_table(const float *logit, const float *delta, const float *mean, int n, int ng,
                       int nstep, float bias, float total, float beta, int form, int *cdf,
                       float *w_out, float *d_out)
{
    float s2 = (float)(1. / sqrt(2.0));
    for (i64 r = 0; r < n; r++) {
        float w[16], d[16];
        float mval = -1e10f, psum = 0;
        for (int i = 0; i < ng; i++) { w[i] = logit[r * ng + i]; if (mval < w[i]) mval = w[i]; }
        for (int i = 0; i < ng; i++) { w[i] = expf(w[i] - mval); psum += w[i]; }
        for (int i = 0; i < ng; i++) w[i] = w[i] / psum;
        for (int i = 0; i < ng; i++) { float t = delta[r * ng + i]; d[i] = t < 0 ? beta : t + beta; }
        if (w_out) for (int i = 0; i < ng; i++) w_out[r * ng + i] = w[i];
        if (d_out) for (int i = 0; i < ng; i++) d_out[r * ng + i] = d[i];
        float c[64];
        c[0] = 0.f;
        c[nstep] = (float)(int)total;
        for (int pt = 1; pt < nstep; pt++) {
            float v = (float)((double)((float)(pt - 1) - bias) + 0.5);
            float ps = 0;
            for (int i = 0; i < ng; i++) {
                float z = s2 * (v - mean[r * ng + i]) / d[i];
                double f = fma(0.5, (double)erff(z), 0.5);
                if (form == 0) ps = (float)fma(f, (double)w[i], (double)ps);   /* batch kernel :146-150: DFMA, one rounding */
                else ps = fmaf((float)f, w[i], ps);                            /* plain kernel :69-72: f stored to float, FFMA */
            }
            c[pt] = (float)(int)((double)(total * ps) + 0.5);
        }
        /* fix-up :83-105 (float arithmetic on integer-valued floats) */
        float fb = 0, mv = 0; int midx = 0;
        for (int i = 0; i < nstep; i++) {
            if (c[i + 1] <= c[i]) fb += 1;
            c[i + 1] += fb;
            if (c[i + 1] - c[i] > mv) { mv = c[i + 1] - c[i]; midx = i; }
        }
        if (fb > 0) for (int i = midx; i < nstep; i++) c[i + 1] -= fb;
        for (int i = 0; i <= nstep; i++) cdf[r * (nstep + 1) + i] = (int)c[i];
    }
}

/* entropy_gmm_cuda.cu:36-69 forward value only: -log(sum_i w_i (Phi((l+.5-mu)/d) - Phi((l-.5-mu)/d)) + 1e-7). */
API void orc_gmm_nll(const float *w, const float *delta, const float *mean, const float *label,
                     float *loss, int n, int ng)
{
    float s2 = (float)(1. / sqrt((double)2.0f));
    for (i64 r = 0; r < n; r++) {
        float sum_p = 0;
        for (int i = 0; i < ng; i++) {
            float xa = (float)((double)label[r] - 0.5 - (double)mean[r * ng + i]);
            float xb = (float)((double)label[r] + 0.5 - (double)mean[r * ng + i]);
            float id = (float)(1. / (double)delta[r * ng + i]);
            float fa = (float)(0.5 + 0.5 * (double)erff(xa * id * s2));
            float fb = (float)(0.5 + 0.5 * (double)erff(xb * id * s2));
            float p = fb - fa;
            sum_p = sum_p + w[r * ng + i] * p;
        }
        loss[r] = (float)(-log((double)sum_p + 0.0000001));
    }
}

/* ------------------------------------------------------------------------------------------------
 * Dense convolution + GDN: fp64-accumulated direct form, the tolerance anchor for the tensor-core
 * kernels (the reference calls cuDNN here: model_zoo_v2.py:41-45 etc., PseudoContextV2.py:207).
 * x (N, Ci, Hi, Wi) already padded, w (Co, Ci, k, k), stride s, no implicit padding.
 * ---------------------------------------------------------------------------------------------- */
API void orc_conv2d(const float *x, const float *w, const float *b, float *y, int N, int Ci, int Hi, int Wi,
                    int Co, int k, int s)
{
    int Ho = (Hi - k) / s + 1, Wo = (Wi - k) / s + 1;
    for (int n = 0; n < N; n++)
        for (int co = 0; co < Co; co++)
            for (int oy = 0; oy < Ho; oy++)
                for (int ox = 0; ox < Wo; ox++) {
                    double acc = b ? (double)b[co] : 0.0;
                    for (int ci = 0; ci < Ci; ci++)
                        for (int ky = 0; ky < k; ky++)
                            for (int kx = 0; kx < k; kx++)
                                acc += (double)x[(((i64)n * Ci + ci) * Hi + (i64)oy * s + ky) * Wi + (i64)ox * s + kx] *
                                       (double)w[(((i64)co * Ci + ci) * k + ky) * k + kx];
                    y[(((i64)n * Co + co) * Ho + oy) * Wo + ox] = (float)acc;
                }
}
