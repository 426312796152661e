"""Process-wide switches of the B200 path."""
import os

# dense convolution implementation of the transforms: 0 (default) = channels-last tiles on the tcgen05/TMEM implicit-GEMM
# kernels (TF32 operands, fp32 accumulate; transforms_nhwc.py), 1 = NCHW fp32 CUDA-core direct form (exact-order on-device
# reference used by the parity tests; model_zoo_v2.pconv).
CONV_IMPL = int(os.environ.get("PCX_CONV_IMPL", "0"))

# wavefront (entropy) loop: 1 (default) = native engine (pcx_wave_encode / pcx_wave_decode: the whole serial loop in one
# native call), 0 = operator-by-operator Python loop shaped like the reference's EntEncoder / EntDecoder.  Same kernels,
# bit-identical CDF tables and bitstreams.
WAVE_IMPL = int(os.environ.get("PCX_WAVE_IMPL", "1"))

# entropy ENCODER inside the native engine: 1 (default) = one-shot (pcx_wave_encode_full: every layer of the context model
# over the whole symbol tensor in one launch, CDF rows emitted in coding order, host coder pipelined behind the device),
# 0 = the stepwise loop (pcx_wave_encode).  Same per-scalar arithmetic, byte-identical bitstreams.
WAVE_ENCODE_FULL = int(os.environ.get("PCX_WAVE_ENCODE_FULL", "1"))
# rows (symbols) per image in one chunk of the one-shot encoder's CDF stream: the host codes chunk i while chunk i+1 is computed
# (8 x 512x1024, 93 K rows per image, entropy encode in ms: one chunk 11.9, 64 K rows per chunk 10.9-11.2, 8 K .. 32 K 10.8-11.9)
WAVE_CHUNK_ROWS = int(os.environ.get("PCX_WAVE_CHUNK_ROWS", str(1 << 16)))

# capture the channels-last analysis / synthesis transforms into CUDA graphs (transforms_nhwc._run_graphed)
CUDA_GRAPHS = os.environ.get("PCX_CUDA_GRAPHS", "1") != "0"
