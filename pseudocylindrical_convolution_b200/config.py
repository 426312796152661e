"""Process-wide switches of the B200 path."""
import os

# dense convolution implementation used by model_zoo_v2.pconv: 0 = tcgen05/TMEM implicit GEMM (TF32 operands,
# fp32 accumulate), 1 = fp32 CUDA-core direct form (exact-order on-device reference).
CONV_IMPL = int(os.environ.get("PCX_CONV_IMPL", "1"))
