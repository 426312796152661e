// pcx_conv_tc.cu - tcgen05 / TMEM implicit-GEMM convolution (TF32 operands, fp32 accumulate).
#include "pcx_common.cuh"

int pcx_conv2d_tc(const pcx_conv_desc *d, const float *d_x, const float *d_w, const float *d_bias, const float *d_slope,
                  const float *d_mul, const float *d_residual, float *d_y, void *stream)
{
    (void)d; (void)d_x; (void)d_w; (void)d_bias; (void)d_slope; (void)d_mul; (void)d_residual; (void)d_y; (void)stream;
    pcx_set_error("tensor-core convolution is not built into this libpcx");
    return PCX_EINVAL;
}

extern "C" long long pcx_conv_pack_weights(const float *d_w, float *d_out, int Co, int Ci, int k, void *stream)
{
    (void)d_w; (void)d_out; (void)stream;
    return (long long)k * k * Co * Ci;
}
