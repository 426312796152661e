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
template <int L>
__global__ void quant_kernel(const float *__restrict__ x, const float *__restrict__ steps, float *__restrict__ val,
                             float *__restrict__ sym, float *__restrict__ count, Bands bands, i64 total, int C, int hw, int W)
{
    extern __shared__ float hist[];          // C * L block-local histogram (only when count != nullptr)
    if (count) {
        for (int i = threadIdx.x; i < C * L; i += blockDim.x) hist[i] = 0.f;
        __syncthreads();
    }
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        int xw = (int)(i % W);
        int c = (int)((i / hw) % C);
        int g = (int)((i / hw / C) % bands.npart);
        if (xw >= bands.wl[g]) {
            val[i] = 0.f;
            if (sym) sym[i] = 0.f;
            continue;
        }
        const float *w = steps + (i64)c * L;
        float v = x[i];
        float tmp = __fsub_rn(v, w[0]);
        int j = 0;
        float out;
        if (tmp < 0.f) {
            out = w[0];
        } else {
            j = 1;
#pragma unroll
            for (; j < L; j++) {
                tmp = __fsub_rn(tmp, w[j]);
                if (tmp < 0.f) break;
            }
            if (j == L) j--;
            if (__fadd_rn(__fadd_rn(tmp, tmp), w[j]) < 0.f) {
                tmp = __fadd_rn(tmp, w[j]);
                j--;
            }
            out = __fsub_rn(v, tmp);
        }
        val[i] = out;
        if (sym) sym[i] = (float)j;
        if (count) atomicAdd(&hist[c * L + j], -1.0f);
    }
    if (count) {
        __syncthreads();
        for (int i = threadIdx.x; i < C * L; i += blockDim.x)
            if (hist[i] != 0.f) atomicAdd(&count[i], hist[i]);
    }
}

// pseudo_dquant_forward_kernel (pseudo_dquant_cuda.cu:34-47)
__global__ void dquant_kernel(const float *__restrict__ sym, const float *__restrict__ centres, float *__restrict__ out,
                              Bands bands, i64 total, int C, int hw, int W, int L)
{
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (i64)gridDim.x * blockDim.x) {
        int xw = (int)(i % W);
        int c = (int)((i / hw) % C);
        int g = (int)((i / hw / C) % bands.npart);
        if (xw >= bands.wl[g]) { out[i] = 0.f; continue; }
        int idx = (int)((double)sym[i] + 0.00001);
        idx = idx < 0 ? 0 : (idx >= L ? L - 1 : idx);     // the reference reads out of bounds here; clamp instead
        out[i] = centres[(i64)c * L + idx];
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
    quant_kernel<8><<<grid_for(total, 256, 8), 256, smem, s>>>(d_x, d_steps, d_val, d_sym, d_count, b, total, C, h * W, W);
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
    dquant_kernel<<<grid_for(total, 256, 8), 256, 0, s>>>(d_sym, d_centres, d_out, b, total, C, h * W, W, L);
    PCX_LAUNCHED();
    return PCX_OK;
}

}  // extern "C"
