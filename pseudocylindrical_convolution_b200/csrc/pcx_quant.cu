// pcx_quant.cu - learned non-uniform scalar quantiser (PseudoQUANTV2 / PseudoDQUANT).
#include "pcx_common.cuh"

namespace {

// pseudo_quant_cal_weight_kernel (extension/pseudo_quant_cuda.cu:37-45): step[c][0] = theta, others exp(theta).
__global__ void quant_steps_kernel(const float *__restrict__ theta, float *__restrict__ steps, int n, int L)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i % L == 0) steps[i] = theta[i];
    else steps[i] = exp(theta[i]);          // same call as the reference: float overload of exp
}

// pseudo_dquant_cal_weight_kernel (extension/pseudo_dquant_cuda.cu:24-31): running sum of exp(theta).
// Kept as the reference's source expression `prev + exp(x)`: nvcc fuses the last multiply of expf with
// the add, and the centres must come out bit-identical.
__global__ void dquant_centres_kernel(const float *__restrict__ input, float *__restrict__ output, int C, int level)
{
    int index = blockIdx.x * blockDim.x + threadIdx.x;
    if (index >= C) return;
    output[index * level] = input[index * level];
    for (int i = 1; i < level; i++) {
        output[index * level + i] = output[index * level + i - 1] + exp(input[index * level + i]);
    }
}

// pseudo_quant_single_gpu_forward_kernel + pseudo_quant_gpu_copy (pseudo_quant_cuda.cu:48-94) in one pass:
// sequential-subtraction search for the bin, nearest-centre tie rule, zero outside the band.
// Branch-free form of the reference's search loop (pseudo_quant_cuda.cu:61-81): the same chain of subtractions
// t_j = t_(j-1) - w[j]; the bin is the FIRST j with t_j < 0 (the reference breaks there), or L-1 with the last remainder.
template <int L>
__device__ __forceinline__ float quant_one(float v, const float *w, int &jout)
{
    float t[L];
    t[0] = __fsub_rn(v, w[0]);
#pragma unroll
    for (int j = 1; j < L; j++) t[j] = __fsub_rn(t[j - 1], w[j]);
    int j = L - 1;
    float tmp = t[L - 1];
#pragma unroll
    for (int k = L - 1; k >= 1; k--)
        if (t[k] < 0.f) { j = k; tmp = t[k]; }
    // nearest-centre tie rule (:77-80)
    const float wj = w[j];                                   // dynamic index over 8 registers -> select chain
    if (__fadd_rn(__fadd_rn(tmp, tmp), wj) < 0.f) {
        tmp = __fadd_rn(tmp, wj);
        j--;
    }
    float out = __fsub_rn(v, tmp);
    if (t[0] < 0.f) { out = w[0]; j = 0; }
    jout = j;
    return out;
}

// One WARP per tensor row (W columns of one channel of one band): the band width and the channel's eight steps are row
// constants, lanes stride over the row in 128-bit vectors (V = 4) or scalars (V = 1, odd widths / unaligned tensors).
template <int L, int V>
__global__ void __launch_bounds__(256) quant_kernel(const float *__restrict__ x, const float *__restrict__ steps,
                                                    float *__restrict__ val, float *__restrict__ sym, float *__restrict__ count,
                                                    Bands bands, i64 nrows, int C, int h, int W)
{
    extern __shared__ float hist[];          // C * L block-local histogram (only when count != nullptr)
    if (count) {
        for (int i = threadIdx.x; i < C * L; i += blockDim.x) hist[i] = 0.f;
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 row = warp0; row < nrows; row += nwarps) {
        const i64 pc = row / h;                        // plane * C + channel
        const int c = (int)(pc % C);
        const int wl = bands.wl[(int)((pc / C) % bands.npart)];
        float w[L];
#pragma unroll
        for (int j = 0; j < L; j++) w[j] = __ldg(steps + (i64)c * L + j);
        const i64 base = row * W;
        for (int x0 = lane * V; x0 < W; x0 += 32 * V) {
            float in[V], ov[V], os[V];
            if (V == 4) {
                if (x0 < wl) {
                    const float4 t = __ldcs(reinterpret_cast<const float4 *>(x + base + x0));
                    in[0] = t.x; in[1 % V] = t.y; in[2 % V] = t.z; in[3 % V] = t.w;
                }
            } else {
                in[0] = x0 < wl ? x[base + x0] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < V; k++) {
                if (x0 + k >= wl) { ov[k] = 0.f; os[k] = 0.f; continue; }
                int j;
                ov[k] = quant_one<L>(in[k], w, j);
                os[k] = (float)j;
                if (count) atomicAdd(&hist[c * L + j], -1.0f);
            }
            if (V == 4) {
                st_cs_f4(val + base + x0, make_float4(ov[0], ov[1 % V], ov[2 % V], ov[3 % V]));
                if (sym) st_cs_f4(sym + base + x0, make_float4(os[0], os[1 % V], os[2 % V], os[3 % V]));
            } else {
                val[base + x0] = ov[0];
                if (sym) sym[base + x0] = os[0];
            }
        }
    }
    if (count) {
        __syncthreads();
        for (int i = threadIdx.x; i < C * L; i += blockDim.x)
            if (hist[i] != 0.f) atomicAdd(&count[i], hist[i]);
    }
}

// pseudo_dquant_forward_kernel (pseudo_dquant_cuda.cu:34-47), same row mapping
template <int V>
__global__ void __launch_bounds__(256) dquant_kernel(const float *__restrict__ sym, const float *__restrict__ centres,
                                                     float *__restrict__ out, Bands bands, i64 nrows, int C, int h, int W, int L)
{
    const int lane = threadIdx.x & 31;
    const i64 warp0 = ((i64)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((i64)gridDim.x * blockDim.x) >> 5;
    for (i64 row = warp0; row < nrows; row += nwarps) {
        const i64 pc = row / h;
        const int c = (int)(pc % C);
        const int wl = bands.wl[(int)((pc / C) % bands.npart)];
        const i64 base = row * W;
        for (int x0 = lane * V; x0 < W; x0 += 32 * V) {
            float in[V], ov[V];
            if (V == 4) {
                if (x0 < wl) {
                    const float4 t = __ldcs(reinterpret_cast<const float4 *>(sym + base + x0));
                    in[0] = t.x; in[1 % V] = t.y; in[2 % V] = t.z; in[3 % V] = t.w;
                }
            } else {
                in[0] = x0 < wl ? sym[base + x0] : 0.f;
            }
#pragma unroll
            for (int k = 0; k < V; k++) {
                if (x0 + k >= wl) { ov[k] = 0.f; continue; }
                int idx = (int)((double)in[k] + 0.00001);
                idx = idx < 0 ? 0 : (idx >= L ? L - 1 : idx);     // the reference reads out of bounds here; clamp instead
                ov[k] = __ldg(centres + (i64)c * L + idx);
            }
            if (V == 4) st_cs_f4(out + base + x0, make_float4(ov[0], ov[1 % V], ov[2 % V], ov[3 % V]));
            else out[base + x0] = ov[0];
        }
    }
}

inline int grid_for(i64 total, int threads, int per_sm)
{
    i64 want = (total + threads - 1) / threads;
    i64 cap = (i64)pcx_sm_count() * per_sm;
    return (int)(want < cap ? (want < 1 ? 1 : want) : cap);
}

}  // namespace

extern "C" {

int pcx_quant_fwd(const float *d_x, const float *d_theta, float *d_steps, float *d_val, float *d_sym, float *d_count,
                  int N, int C, int h, int W, int npart, int L, const int *wl, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_x && d_theta && d_steps && d_val, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W > 0, "bad shape");
    PCX_REQUIRE(L == 8, "only the reference's 8-level quantiser is built (bin_num=%d)", L);
    PCX_REQUIRE((size_t)C * L * sizeof(float) <= 48 * 1024, "channel count %d too large for the histogram", C);
    cudaStream_t s = (cudaStream_t)stream;
    quant_steps_kernel<<<ceil_div((i64)C * L, 256), 256, 0, s>>>(d_theta, d_steps, C * L, L);
    PCX_LAUNCHED();
    if (d_count) PCX_CUDA(cudaMemsetAsync(d_count, 0, (size_t)C * L * sizeof(float), s));   // caffe_gpu_set(...,0,count) :172
    i64 total = (i64)N * npart * C * h * W;
    size_t smem = d_count ? (size_t)C * L * sizeof(float) : 0;
    const bool v4 = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_val) |
                                    reinterpret_cast<uintptr_t>(d_sym)) & 15) == 0;
    const i64 nrows = total / W;
    const int grid = grid_for(nrows * 32, 256, 8);
    if (v4) quant_kernel<8, 4><<<grid, 256, smem, s>>>(d_x, d_steps, d_val, d_sym, d_count, b, nrows, C, h, W);
    else quant_kernel<8, 1><<<grid, 256, smem, s>>>(d_x, d_steps, d_val, d_sym, d_count, b, nrows, C, h, W);
    PCX_LAUNCHED();
    return PCX_OK;
}

int pcx_dquant_fwd(const float *d_sym, const float *d_theta, float *d_centres, float *d_out, int N, int C, int h, int W,
                   int npart, int L, const int *wl, void *stream)
{
    Bands b;
    PCX_REQUIRE(make_bands(b, wl, npart) == 0, "bad band description");
    PCX_REQUIRE(d_sym && d_theta && d_centres && d_out, "null pointer");
    PCX_REQUIRE(N > 0 && C > 0 && h > 0 && W > 0 && L > 0, "bad shape");
    cudaStream_t s = (cudaStream_t)stream;
    dquant_centres_kernel<<<ceil_div(C, 128), 128, 0, s>>>(d_theta, d_centres, C, L);
    PCX_LAUNCHED();
    i64 total = (i64)N * npart * C * h * W;
    const bool v4 = W % 4 == 0 && ((reinterpret_cast<uintptr_t>(d_sym) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    const i64 nrows = total / W;
    const int grid = grid_for(nrows * 32, 256, 8);
    if (v4) dquant_kernel<4><<<grid, 256, 0, s>>>(d_sym, d_centres, d_out, b, nrows, C, h, W, L);
    else dquant_kernel<1><<<grid, 256, 0, s>>>(d_sym, d_centres, d_out, b, nrows, C, h, W, L);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
