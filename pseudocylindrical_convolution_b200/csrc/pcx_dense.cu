// pcx_dense.cu - dense contractions of the analysis / synthesis transforms on CUDA cores (fp32, fixed
// accumulation order): the direct convolution used as the on-device exact-order reference for the tcgen05
// implicit-GEMM kernels (pcx_conv_tc.cu), and GDN / IGDN.
#include "pcx_common.cuh"

int pcx_conv2d_tc(const pcx_conv_desc *d, const float *d_x, const float *d_w, const float *d_bias, const float *d_slope,
                  const float *d_mul, const float *d_residual, float *d_y, void *stream);

namespace {

constexpr int PIX = 128;     // output pixels per CTA (one per thread, consecutive along x)
constexpr int COT = 32;      // output channels per CTA
constexpr int CIC = 8;       // input channels staged per weight chunk

struct ConvGeom {
    pcx_conv_desc d;
};

// y = fill( residual + mul * act(conv(x) + bias) ), accumulation order: ci ascending, ky, kx; one FFMA chain.
template <int K>
__global__ void __launch_bounds__(PIX) conv_direct_kernel(ConvGeom G, const float *__restrict__ x, const float *__restrict__ w,
                                                          const float *__restrict__ bias, const float *__restrict__ slope,
                                                          const float *__restrict__ mul, const float *__restrict__ residual,
                                                          float *__restrict__ y)
{
    const pcx_conv_desc &d = G.d;
    __shared__ __align__(16) float ws[CIC][K * K][COT];
    const int xt = (d.Wo + PIX - 1) / PIX;
    const int cot = (d.Co + COT - 1) / COT;
    i64 bid = blockIdx.x;
    const int bx = (int)(bid % xt); bid /= xt;
    const int oy = (int)(bid % d.Ho); bid /= d.Ho;
    const int ct = (int)(bid % cot); bid /= cot;
    const i64 plane = bid;                       // image * npart + band
    const int g = (int)(plane % d.npart);
    const int co0 = ct * COT;
    const int ox = bx * PIX + threadIdx.x;
    const int wl = d.wl_out[g];
    const bool live = ox < d.Wo && ox < wl;

    float acc[COT];
#pragma unroll
    for (int i = 0; i < COT; i++) acc[i] = 0.f;

    // CTA-uniform early out: nothing valid in this pixel tile
    const bool tile_live = bx * PIX < wl;
    if (tile_live) {
        const float *xp = x + (plane * d.Ci * d.Hi + (i64)oy * d.stride) * d.in_pitch + (i64)ox * d.stride;
        for (int c0 = 0; c0 < d.Ci; c0 += CIC) {
            __syncthreads();
            for (int i = threadIdx.x; i < CIC * K * K * COT; i += PIX) {
                int co = i % COT, tap = (i / COT) % (K * K), ci = i / COT / (K * K);
                float v = 0.f;
                if (c0 + ci < d.Ci && co0 + co < d.Co) v = w[((i64)(co0 + co) * d.Ci + c0 + ci) * (K * K) + tap];
                ws[ci][tap][co] = v;
            }
            __syncthreads();
            if (live) {
                const int nci = min(CIC, d.Ci - c0);
                for (int ci = 0; ci < nci; ci++) {
                    const float *xc = xp + (i64)(c0 + ci) * d.Hi * d.in_pitch;
#pragma unroll
                    for (int ky = 0; ky < K; ky++)
#pragma unroll
                        for (int kx = 0; kx < K; kx++) {
                            float xv = __ldg(xc + (i64)ky * d.in_pitch + kx);
                            const float4 *wv = reinterpret_cast<const float4 *>(&ws[ci][ky * K + kx][0]);
#pragma unroll
                            for (int q = 0; q < COT / 4; q++) {
                                float4 t = wv[q];
                                acc[q * 4 + 0] = __fmaf_rn(xv, t.x, acc[q * 4 + 0]);
                                acc[q * 4 + 1] = __fmaf_rn(xv, t.y, acc[q * 4 + 1]);
                                acc[q * 4 + 2] = __fmaf_rn(xv, t.z, acc[q * 4 + 2]);
                                acc[q * 4 + 3] = __fmaf_rn(xv, t.w, acc[q * 4 + 3]);
                            }
                        }
                }
            }
        }
    }
    if (ox >= d.Wo) return;
#pragma unroll
    for (int i = 0; i < COT; i++) {
        const int co = co0 + i;
        if (co >= d.Co) break;
        float v = 0.f;
        if (live) {
            v = acc[i];
            if (bias) v = __fadd_rn(v, bias[co]);
            if (d.act == 1) { if (v < 0.f) v = __fmul_rn(v, slope[co]); }
            else if (d.act == 2) v = 1.0f / (1.0f + expf(-v));
            if (mul || residual) {
                i64 a = ((plane * d.Co + co) * d.aux_rows + oy + d.aux_y0) * (i64)d.aux_pitch + ox + d.aux_x0;
                if (mul) v = __fmul_rn(v, mul[a]);
                if (residual) v = __fadd_rn(residual[a], v);
            }
        }
        y[((plane * d.Co + co) * d.out_rows + oy + d.out_y0) * (i64)d.out_pitch + ox + d.out_x0] = v;
    }
}

// ------------------------------------------------------------------------------------------------ GDN
// LowerBound + reparametrisation (PCONV_operator/PseudoContextV2.py:196-203, GDN.py:6-22):
// beta' = max(beta, beta_bound)^2 - pedestal, gamma' likewise with gamma_bound.
__global__ void gdn_params_kernel(const float *__restrict__ beta, const float *__restrict__ gamma, float *__restrict__ beta_eff,
                                  float *__restrict__ gamma_eff, int C, float beta_bound, float gamma_bound, float pedestal)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < C) {
        float b = fmaxf(beta[i], beta_bound);
        beta_eff[i] = __fsub_rn(__fmul_rn(b, b), pedestal);
    }
    if (i < C * C) {
        float gm = fmaxf(gamma[i], gamma_bound);
        gamma_eff[i] = __fsub_rn(__fmul_rn(gm, gm), pedestal);
    }
}

constexpr int GPIX = 64;      // pixels per CTA
constexpr int GTHREADS = 256; // 4 channel quarters x 64 pixels

// y[c] = x[c] / sqrt(beta'[c] + sum_j gamma'[c][j] x[j]^2)   (inverse: multiply), 0 outside the band.
// gamma' is staged transposed in shared memory ([j][c], broadcast float4 reads), x^2 as [j][pixel].
template <int C>
__global__ void __launch_bounds__(GTHREADS) gdn_kernel(const float *__restrict__ x, const float *__restrict__ beta_eff,
                                                       const float *__restrict__ gamma_eff, const float *__restrict__ residual,
                                                       float *__restrict__ y, Bands bands, int h, int W, int inverse, i64 ntiles)
{
    extern __shared__ __align__(16) float sm[];
    float *gT = sm;                    // [C][C] : gT[j*C + c] = gamma'[c][j], loaded once per (persistent) CTA
    float *xs = sm + C * C;            // [C][GPIX] squares
    constexpr int CQ = C / 4;          // output channels per thread
    const int xt = (W + GPIX - 1) / GPIX;
    const int px = threadIdx.x % GPIX, cq = threadIdx.x / GPIX;
    const i64 plane_stride = (i64)h * W;
    for (int i = threadIdx.x; i < C * C; i += GTHREADS) {
        int c = i / C, j = i % C;
        gT[j * C + c] = gamma_eff[i];
    }
    float b0[CQ];
#pragma unroll
    for (int i = 0; i < CQ; i++) b0[i] = beta_eff[cq * CQ + i];

    for (i64 t = blockIdx.x; t < ntiles; t += gridDim.x) {
        i64 bid = t;
        const int bx = (int)(bid % xt); bid /= xt;
        const int row = (int)(bid % h); bid /= h;
        const i64 tile = bid;
        const int g = (int)(tile % bands.npart);
        const int wl = bands.wl[g];
        const int ox = bx * GPIX + px;
        const float *xb = x + tile * C * plane_stride + (i64)row * W;
        float *yb = y + tile * C * plane_stride + (i64)row * W;
        if (bx * GPIX >= wl) {                       // whole pixel tile outside the band: zeros
            for (int c = cq; c < C; c += 4)
                if (ox < W) yb[(i64)c * plane_stride + ox] = 0.f;
            continue;
        }
        __syncthreads();                             // previous tile's xs fully consumed (and gT visible)
        for (int i = threadIdx.x; i < C * GPIX; i += GTHREADS) {
            int j = i / GPIX, p = i % GPIX;
            int xx = bx * GPIX + p;
            float v = (xx < wl) ? xb[(i64)j * plane_stride + xx] : 0.f;
            xs[j * GPIX + p] = __fmul_rn(v, v);
        }
        __syncthreads();
        float acc[CQ];
#pragma unroll
        for (int i = 0; i < CQ; i++) acc[i] = b0[i];
        for (int j = 0; j < C; j++) {
            float sq = xs[j * GPIX + px];
            const float4 *gv = reinterpret_cast<const float4 *>(gT + j * C + cq * CQ);
#pragma unroll
            for (int q = 0; q < CQ / 4; q++) {
                float4 tt = gv[q];
                acc[q * 4 + 0] = __fmaf_rn(sq, tt.x, acc[q * 4 + 0]);
                acc[q * 4 + 1] = __fmaf_rn(sq, tt.y, acc[q * 4 + 1]);
                acc[q * 4 + 2] = __fmaf_rn(sq, tt.z, acc[q * 4 + 2]);
                acc[q * 4 + 3] = __fmaf_rn(sq, tt.w, acc[q * 4 + 3]);
            }
        }
        if (ox >= W) continue;
#pragma unroll
        for (int i = 0; i < CQ; i++) {
            const int c = cq * CQ + i;
            float out = 0.f;
            if (ox < wl) {
                float v = xb[(i64)c * plane_stride + ox];
                float nrm = sqrtf(acc[i]);
                out = inverse ? __fmul_rn(v, nrm) : __fdiv_rn(v, nrm);
                if (residual) out = __fadd_rn(residual[tile * C * plane_stride + (i64)row * W + (i64)c * plane_stride + ox], out);
            }
            yb[(i64)c * plane_stride + ox] = out;
        }
    }
}

}  // namespace

extern "C" {

int pcx_conv2d_fwd(const pcx_conv_desc *desc, const float *d_x, const float *d_w, const float *d_bias, const float *d_slope,
                   const float *d_mul, const float *d_residual, float *d_y, void *stream)
{
    PCX_REQUIRE(desc && d_x && d_w && d_y, "null pointer");
    const pcx_conv_desc &d = *desc;
    PCX_REQUIRE(d.N > 0 && d.npart > 0 && d.npart <= PCX_MAX_PART, "bad batch N=%d npart=%d", d.N, d.npart);
    PCX_REQUIRE(d.k == 1 || d.k == 3, "kernel size %d not supported (1 or 3)", d.k);
    PCX_REQUIRE(d.stride == 1 || d.stride == 2, "stride %d not supported (1 or 2)", d.stride);
    PCX_REQUIRE(d.Ci > 0 && d.Co > 0 && d.Ho > 0 && d.Wo > 0, "bad extent");
    PCX_REQUIRE((d.Ho - 1) * d.stride + d.k <= d.Hi, "output rows %d read past the %d input rows", d.Ho, d.Hi);
    PCX_REQUIRE((d.Wo - 1) * d.stride + d.k <= d.in_pitch, "output columns %d read past the input pitch %d", d.Wo, d.in_pitch);
    const int up = d.impl == 3 ? 2 : 1;      // impl 3 writes the result depth-to-space (2 Ho x 2 Wo pixels of Co / 4 channels)
    PCX_REQUIRE(d.out_y0 >= 0 && d.out_x0 >= 0 && d.out_y0 + up * d.Ho <= d.out_rows && d.out_x0 + up * d.Wo <= d.out_pitch, "output window outside the output plane");
    PCX_REQUIRE(d.impl != 3 || (d.Co % 4 == 0 && !d_mul && !d_residual), "impl 3 (fused depth-to-space) needs Co %% 4 == 0 and no gate / residual operand");
    PCX_REQUIRE(d.act >= 0 && d.act <= 4, "act %d", d.act);
    PCX_REQUIRE(d.act <= 2 || d.impl != 1, "act %d (rsqrt / sqrt) is implemented by the tensor-core path only", d.act);
    PCX_REQUIRE(d.in_plane_rows == 0 || (d.in_plane_rows >= d.Hi && d.impl != 1), "in_plane_rows %d (tensor-core path only, >= Hi)", d.in_plane_rows);
    PCX_REQUIRE(d.act != 1 || d_slope, "PReLU needs slopes");
    PCX_REQUIRE(d.square_input == 0 || d.square_input == 1, "square_input %d", d.square_input);
    PCX_REQUIRE(d.square_input == 0 || d.impl != 1, "square_input is implemented by the tensor-core path only");
    if (d_mul || d_residual)
        PCX_REQUIRE(d.aux_y0 >= 0 && d.aux_x0 >= 0 && d.aux_y0 + d.Ho <= d.aux_rows && d.aux_x0 + d.Wo <= d.aux_pitch, "aux window outside the aux plane");
    if (d.impl == 0 || d.impl == 2 || d.impl == 3) return pcx_conv2d_tc(desc, d_x, d_w, d_bias, d_slope, d_mul, d_residual, d_y, stream);
    PCX_REQUIRE(d.impl == 1, "impl %d", d.impl);
    ConvGeom G;
    G.d = d;
    i64 blocks = (i64)((d.Wo + PIX - 1) / PIX) * d.Ho * ((d.Co + COT - 1) / COT) * d.N * d.npart;
    PCX_REQUIRE(blocks < (1ll << 31), "grid too large");
    cudaStream_t s = (cudaStream_t)stream;
    if (d.k == 1) conv_direct_kernel<1><<<(unsigned)blocks, PIX, 0, s>>>(G, d_x, d_w, d_bias, d_slope, d_mul, d_residual, d_y);
    else conv_direct_kernel<3><<<(unsigned)blocks, PIX, 0, s>>>(G, d_x, d_w, d_bias, d_slope, d_mul, d_residual, d_y);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_gdn_params(const float *d_beta, const float *d_gamma, float *d_beta_eff, float *d_gamma_eff, int C, float beta_min,
                   float reparam_offset, void *stream)
{
    PCX_REQUIRE(d_beta && d_gamma && d_beta_eff && d_gamma_eff && C > 0, "bad arguments");
    // PseudoContextV2.py:159-163 evaluates these in float32 tensors
    float pedestal = reparam_offset * reparam_offset;
    float beta_bound = sqrtf(beta_min + pedestal);
    gdn_params_kernel<<<ceil_div((i64)C * C, 256), 256, 0, (cudaStream_t)stream>>>(d_beta, d_gamma, d_beta_eff, d_gamma_eff, C,
                                                                                   beta_bound, reparam_offset, pedestal);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_gdn_fwd(const float *d_x, const float *d_beta_eff, const float *d_gamma_eff, const float *d_residual, float *d_y, int N,
                int C, int h, int W, int npart, const int *wl, int inverse, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_x && d_beta_eff && d_gamma_eff && d_y, "null pointer");
    PCX_REQUIRE(C == 192, "GDN is built for the codec's 192 channels (got %d)", C);
    i64 ntiles = (i64)((W + GPIX - 1) / GPIX) * h * N * npart;
    i64 blocks = ntiles < pcx_sm_count() ? ntiles : pcx_sm_count();
    size_t smem = (size_t)(192 * 192 + 192 * GPIX) * sizeof(float);
    static PcxDeviceOnce once;
    PCX_ONCE_PER_DEVICE(once) PCX_CUDA(cudaFuncSetAttribute(gdn_kernel<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gdn_kernel<192><<<(unsigned)blocks, GTHREADS, smem, (cudaStream_t)stream>>>(d_x, d_beta_eff, d_gamma_eff, d_residual, d_y, b, h, W, inverse, ntiles);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
