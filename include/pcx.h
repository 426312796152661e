/*
 * pcx.h - C ABI of libpcx.so, the B200 (sm_100a) implementation of the pseudocylindrical codec hot path.
 *
 * This is the drop-in boundary: the reference binds its CUDA operators to Python
 * through pybind11 classes in extension/main.cpp:4-137 (module `PCONV`) and its arithmetic coder through
 * coder/python.cpp:63-72 (module `coder`).  Every entry point below replaces the compute method of one of
 * those classes; the comment above each names the reference interface (file:line under /root/reference).
 * The stateful parts of the reference objects (cached output buffers, the wavefront step counter `pidx_`,
 * the per-shape table caches) live in the host-side mirror `pseudocylindrical_convolution_b200/PCONV.py`,
 * which calls these functions through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch/ATen types.  `d_` = device pointer, `h_` = host pointer.
 *   - all activation tensors are fp32, NCHW, contiguous unless a pitch argument says otherwise; the batch
 *     dimension of a tiled tensor is  image * npart + band  (extension/sphere_slice_cuda.cu:95-97).
 *   - indices are 64-bit inside the kernels (the reference's int32/float-stored offsets overflow above
 *     512x1024x192).
 *   - `wl` = host array of npart band widths (pcx_band_widths); npart <= PCX_MAX_PART.
 *   - `stream` is a cudaStream_t passed as void*; launches are asynchronous on it.
 *   - return value: 0 on success, a negative PCX_E* code otherwise; pcx_last_error() gives the text.
 *     There is no CPU fallback: without a usable sm_100 device every compute entry point fails.
 */
#ifndef PCX_H
#define PCX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* libpcx.so is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define PCX_MAX_PART 32
#define PCX_MAX_GAUSS 16
typedef struct pcx_coder pcx_coder;      /* host arithmetic coder handle, see the end of this header */

enum {
    PCX_OK = 0,
    PCX_EINVAL = -1,   /* bad argument (the reference asserts: math_cuda.cu:225, pseudo_context_cuda.cu:38, entropy_gmm_table_cuda.cu:13) */
    PCX_ECUDA = -2,    /* CUDA runtime / launch error (the reference only printf()s: caffe_cuda_macro.h:21-26) */
    PCX_ENODEV = -3,   /* no sm_100 device */
    PCX_ECODER = -4,   /* arithmetic coder state error (the reference throws const char*: ArithmeticCoder.cpp:36-48) */
    PCX_EIO = -5
};

/* ---- library ------------------------------------------------------------------------------------ */
int pcx_abi_version(void);   /* 2: pcx_conv_desc carries square_input */
const char *pcx_last_error(void);
/* number of kernels launched by this library since load (bench.py's gpu_launches) */
long long pcx_launch_count(void);
int pcx_device_check(int device, int *sm_count, int *cc);

/* ---- geometry ------------------------------------------------------------------------------------
 * replaces sphere_cal_npart_hw_v2 / _v3 (extension/math_cuda.cu:177-253) used by every op's reshape(). */
int pcx_band_widths(const float *weight, int npart, int H, int W, int *h_wl);

/* ---- gather tables (device kernels, double arithmetic identical to the reference kernels) ----------
 * slice:  init_slice_param_kernel   (extension/sphere_slice_cuda.cu:13-32)
 * uslice: init_uslice_param_kernel  (extension/sphere_uslice_cuda.cu:13-30)
 * d_src  int32 [npart][W]   integer source column of the second tap
 * d_wt   float [npart][W][4] Catmull-Rom weights */
int pcx_slice_table(const int *wl, int npart, int W, int *d_src, float *d_wt, void *stream);
int pcx_uslice_table(const int *wl, int npart, int W, int *d_src, float *d_wt, void *stream);
/* halo tables: mode 0 = pseudo_context_forward_kernel (extension/pseudo_context_cuda.cu:51-104),
 *              mode 1 = entropy_context_kernel (extension/entropy_context_cuda.cu:105-165) and
 *                       pseudo_entropy_context_forward_kernel_v1 (pseudo_entropy_context_cuda.cu:111-170),
 *              mode 2 = pseudo_entropy_context_forward_kernel_v0 (pseudo_entropy_context_cuda.cu:50-109).
 * d_band/d_row int32 [npart][2][pad]; d_col int32 / d_tw float [npart][2][pad][W]. */
int pcx_halo_table(const int *wl, int npart, int h, int W, int pad, int mode,
                   int *d_band, int *d_row, int *d_col, float *d_tw, void *stream);

/* ---- tile pipeline --------------------------------------------------------------------------------
 * SphereSliceOp.forward   (main.cpp:37-41 -> sphere_slice_opt::forward_cuda, sphere_slice_cuda.cu:119-144)
 * in (N,C,H,W) -> out (N*npart, C, H/npart + 2*pad, W + 2*pad); only the interior is written when pad>0. */
int pcx_slice_fwd(const float *d_in, float *d_out, int N, int C, int H, int W, int npart, const int *wl,
                  const int *d_src, const float *d_wt, int pad, void *stream);
/* SphereUsliceOp.forward  (main.cpp:43-47 -> sphere_uslice_cuda.cu:101-125) */
int pcx_uslice_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, const int *wl,
                   const int *d_src, const float *d_wt, int pad, void *stream);
/* PseudoPadOp.forward     (main.cpp:97-101 -> pseudo_pad_opt::forward_cuda, pseudo_pad.cu:98-124; three
 * kernels fused into one pass).  in (N*npart,C,h,W) -> out rows of `out_pitch` floats (>= W+2*pad; pass
 * W+2*pad for the reference's contiguous layout), out plane = (h+2*pad) rows. */
int pcx_pad_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int pad, const int *wl,
                const int *d_band, const int *d_row, const int *d_col, const float *d_tw,
                int out_pitch, void *stream);
/* PseudoEntropyPadOp.forward (main.cpp:115-119 -> pseudo_entropy_pad_cuda.cu:107-133): causal halo,
 * left pad 0, right pad wrap; table from pcx_halo_table(mode 1 or 2). */
int pcx_entropy_pad_fwd(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int pad,
                        const int *wl, const int *d_band, const int *d_row, const int *d_col,
                        const float *d_tw, void *stream);
/* In-place halo refresh of an already padded, pitched tile buffer (B200-native replacement of the
 * pad copy inside the fused transforms; same values as pcx_pad_fwd). */
int pcx_halo_fill(float *d_buf, int N, int C, int h, int W, int npart, int pad, const int *wl,
                  const int *d_band, const int *d_row, const int *d_col, const float *d_tw,
                  int pitch, void *stream);
/* B200-native fused forms used by the tensor-core transforms (channels-last tiles):
 * SphereSliceOp.forward + PseudoPadOp.forward in one pass (sphere_slice_cuda.cu:87-116, pseudo_pad.cu:39-96):
 * in (N,C,H,W) NCHW -> out [N*npart][H/npart + 2*pad][out_pitch][C]; d_src/d_wt from pcx_slice_table, halo tables
 * from pcx_halo_table(mode 0) (ignored when pad == 0).  Values are bit-identical to slice followed by pad.
 * zero_invalid != 0 also writes the zeros of the columns >= wl + 2*pad (skip it for pre-zeroed buffers). */
int pcx_slice_pad_nhwc(const float *d_in, float *d_out, int N, int C, int H, int W, int npart, int pad, const int *wl,
                       const int *d_src, const float *d_wt, const int *d_band, const int *d_row, const int *d_col,
                       const float *d_tw, int out_pitch, int zero_invalid, void *stream);
/* SphereUsliceOp.forward (sphere_uslice_cuda.cu:73-99) reading channels-last tiles
 * [N*npart][in_rows][in_pitch][C] whose band data start at (in_y0, in_x0); out (N,C,h*npart,W) NCHW. */
int pcx_uslice_nhwc(const float *d_in, float *d_out, int N, int C, int h, int W, int npart, int in_rows, int in_pitch,
                    int in_y0, int in_x0, const int *wl, const int *d_src, const float *d_wt, void *stream);
/* B200-native channels-last companions of the tensor-core transforms:
 * PseudoPadOp.forward as an in-place halo refresh of a channels-last tile buffer [N*npart][rows][pitch][C] whose band
 * interiors start at (y0, x0) (y0, x0 >= pad); tables from pcx_halo_table(mode 0).  Same values as pcx_pad_fwd. */
int pcx_halo_fill_nhwc(float *d_buf, int N, int C, int h, int W, int npart, int pad, const int *wl, const int *d_band,
                       const int *d_row, const int *d_col, const float *d_tw, int rows, int pitch, int y0, int x0, void *stream);
/* DtowOp.forward (d2w, stride 2; dtow_cuda.cu:38-55) on channels-last tiles: in [planes][in_rows][in_pitch][4*Co] window at
 * (in_y0, in_x0) of extent h x W  ->  out [planes][out_rows][out_pitch][Co] window at (out_y0, out_x0) of extent 2h x 2W. */
int pcx_dtow_nhwc(const float *d_in, float *d_out, int planes, int Co, int h, int W, int in_rows, int in_pitch, int in_y0,
                  int in_x0, int out_rows, int out_pitch, int out_y0, int out_x0, void *stream);
/* y = x * x elementwise (the x^2 operand of the GDN contraction, PseudoContextV2.py:207); n multiple of 4. */
int pcx_square(const float *d_in, float *d_out, long long n, void *stream);
/* PseudoFillOp.forward    (main.cpp:103-107 -> pseudo_fill_cuda.cu:46-61), in place. */
int pcx_fill(float *d_data, int N, int C, int Hh, int Ww, int npart, int pad, int trim, const int *wl,
             float fvalue, void *stream);
/* DtowOp.forward          (main.cpp:12-16 -> dtow_cuda.cu:77-102) */
int pcx_dtow(const float *d_in, float *d_out, int N, int C, int H, int W, int stride, int d2w, void *stream);

/* ---- quantiser -------------------------------------------------------------------------------------
 * PseudoQuantOp.forward (main.cpp:121-125 -> pseudo_quant_opt::quant_forward_cuda, pseudo_quant_cuda.cu:157-194)
 * d_theta (C,L) learned parameters; d_steps (C,L) scratch (expanded step table, :37-45);
 * d_val = dequantised value, d_sym = symbol as float (may be NULL), d_count (C,L) histogram side effect
 * (may be NULL; accumulates -1 per symbol like :65,:83). */
int pcx_quant_fwd(const float *d_x, const float *d_theta, float *d_steps, float *d_val, float *d_sym,
                  float *d_count, int N, int C, int h, int W, int npart, int L, const int *wl, void *stream);
/* PseudoDQuantOp.forward (main.cpp:127-130 -> pseudo_dquant_cuda.cu:49-70); d_centres (C,L) scratch. */
int pcx_dquant_fwd(const float *d_sym, const float *d_theta, float *d_centres, float *d_out,
                   int N, int C, int h, int W, int npart, int L, const int *wl, void *stream);

/* ---- dense pseudocylindrical convolution -------------------------------------------------------------
 * replaces PseudoPadV2 -> nn.Conv2d(padding=0) [cuDNN] -> PReLU/... -> PseudoFillV2 of model_zoo_v2.py:36-211.
 * Input is an already padded tile tensor (N*npart, Ci, Hi, in_pitch) whose valid data start at column 0;
 * output pixel (y,x) reads input rows y*stride .. y*stride+k-1, columns x*stride .. x*stride+k-1.
 * Output (N*npart, Co, Ho, out_pitch) is written at row offset out_y0 / column offset out_x0 of each plane
 * (so a layer can write straight into the interior of the next layer's padded buffer).
 * Epilogue, in this order:  +bias ; act ; +residual ; zero columns >= wl_out[band] (PseudoFillV2).
 *   act: 0 none, 1 PReLU(d_slope per channel), 2 sigmoid, 3 1/sqrt(.), 4 sqrt(.)  (3/4 with d_mul = x and a 1x1 conv of x^2
 *        by gamma' + beta' are GDN / IGDN, PseudoContextV2.py:186-216; tensor-core impl only).
 * impl: 0 = tcgen05/TMEM implicit GEMM (TF32 operands, fp32 accumulate); tensors are NHWC
 *           (x [plane][Hi][in_pitch][Ci], y [plane][out_rows][out_pitch][Co], aux likewise), Ci % 32 == 0,
 *           Co <= 16 or Co % 96 == 0;
 *       1 = fp32 CUDA-core direct form on NCHW tensors (exact-order reference used to validate impl 0);
 *       2 = as 0, but d_w is already packed by pcx_conv_pack_weights (no per-call repack);
 *       3 = as 2 with Dtow(stride 2, d2w) fused into the store (ResidualBlockUp / the decoder's last layer run
 *           conv -> Dtow, model_zoo_v2.py:153-175, dtow_cuda.cu:38-55): d_w packed by pcx_conv_pack_weights_d2w,
 *           y is [plane][out_rows][out_pitch][Co/4] and receives 2 Ho x 2 Wo pixels per plane at (out_y0, out_x0):
 *           channel 4c + 2dy + dx of conv pixel (y,x) -> channel c of pixel (2y+dy, 2x+dx).  No d_mul / d_residual;
 *           Co / 4 must be 96 or 192.  wl_out stays in conv-pixel units. */
typedef struct pcx_conv_desc {
    int N, npart;            /* images, bands per image */
    int Ci, Hi, in_pitch;    /* input view: Hi rows of in_pitch pixels reachable from d_x */
    int Co, Ho, Wo;          /* output extent computed per plane */
    int out_rows, out_pitch; /* output plane geometry (rows x pitch) */
    int out_y0, out_x0;      /* where (0,0) of the result lands inside the output plane */
    int k, stride;           /* 1 or 3; 1 or 2 */
    int act;                 /* 0 none, 1 PReLU, 2 sigmoid, 3 rsqrt, 4 sqrt */
    int impl;                /* 0 tensor core, 1 fp32 direct, 2 tensor core with pre-packed weights */
    int aux_rows, aux_pitch; /* geometry shared by the optional d_mul / d_residual planes (Co channels) */
    int aux_y0, aux_x0;
    int in_plane_rows;       /* physical rows between consecutive input planes (0 = Hi); lets d_x point INTO a padded buffer */
    int wl_out[PCX_MAX_PART];/* valid output width per band (columns >= wl_out are written as 0) */
    int square_input;        /* 1: the convolution reads x * x (GDN / IGDN: beta' + gamma' x^2, PseudoContextV2.py:186-216) - squared on
                              * chip, x^2 never exists in HBM; tensor-core path, k = 1, act 3 or 4 */
} pcx_conv_desc;
/* y = fill( residual + mul * act(conv(x) + bias) );  d_bias, d_slope, d_mul, d_residual may be NULL.
 * (AttentionBlock: x + t * sigmoid(conv1x1(.)), model_zoo_v2.py:73-76; ResidualBlock*: x + y, :49-53, :89-93) */
int pcx_conv2d_fwd(const pcx_conv_desc *desc, const float *d_x, const float *d_w, const float *d_bias,
                   const float *d_slope, const float *d_mul, const float *d_residual, float *d_y, void *stream);
/* Repack OIHW fp32 weights into the tap-major layout the tensor-core kernel streams with TMA:
 * d_out [k*k][Co_pad][Ci_pad]. Returns the element count needed when d_out == NULL. */
long long pcx_conv_pack_weights(const float *d_w, float *d_out, int Co, int Ci, int k, void *stream);
/* Same layout with the output channels permuted for impl 3 (GEMM column q * Co/4 + c <- channel 4c + q). */
long long pcx_conv_pack_weights_d2w(const float *d_w, float *d_out, int Co, int Ci, int k, void *stream);
/* PseudoGDNV2.forward (PCONV_operator/PseudoContextV2.py:186-216): y = x / sqrt(beta' + gamma' x^2)
 * (inverse: x * sqrt(..)), invalid columns -> 0.  beta'/gamma' are the reparametrised values computed by
 * pcx_gdn_params from the raw parameters (LowerBound, GDN.py:6-22). */
int pcx_gdn_params(const float *d_beta, const float *d_gamma, float *d_beta_eff, float *d_gamma_eff,
                   int C, float beta_min, float reparam_offset, void *stream);
/* d_residual (same shape as x, may be NULL): y = fill(residual + gdn(x)) - the block tail `trim(t + y)` of
 * ResidualBlockDown / ResidualBlockUp (model_zoo_v2.py:107-114, :166-175) fused into the same pass. */
int pcx_gdn_fwd(const float *d_x, const float *d_beta_eff, const float *d_gamma_eff, const float *d_residual,
                float *d_y, int N, int C, int h, int W, int npart, const int *wl, int inverse, void *stream);

/* ---- wavefront context model ------------------------------------------------------------------------
 * EntropyContextOp (main.cpp:55-59): entropy_context::reshape_hw (entropy_context_cuda.cu:13-45) builds the
 * anti-diagonal order on the host; these fill HOST arrays (h_order: Hf*W ints, h_start: Hf+W ints). */
int pcx_ctx_order(const int *wl, int npart, int h, int W, int *h_order, int *h_start);
/* per-plane halo / right-wrap work lists (entropy_context_cuda.cu:64-103, :187-204), deterministic order.
 * h_band/h_col/h_tw: host copies of the mode-1 halo table.  h_items: int4 records, h_pstart: Hf+W+pad ints.
 * Pass h_items == NULL to get the count. */
int pcx_ctx_pad_items(const int *wl, int npart, int h, int W, int pad, const int *h_band, const int *h_col,
                      const float *h_tw, int *h_items, int *h_pstart);
/* EntropyCtxPadRun2Op.forward (main.cpp:61-66 -> entropy_ctx_pad_run2_cuda.cu:86-117), in place.
 * psum = step index after the input-layer lag (the caller subtracts 1 when `input` is set, :93-95). */
int pcx_ctx_pad_step(float *d_buf, int nrep, int npart, int G, int cpn, int h, int W, int pad, int psum,
                     const int *wl, const int *d_band, const int *d_row, const int *d_col, const float *d_tw,
                     const int *d_items, const int *h_pstart, void *stream);
/* EntropyConv2Op.forward / forward_act / forward_batch / forward_act_batch (main.cpp:81-88 ->
 * entropy_conv_cuda_v2.cu:292-323, :381-414).  d_act == NULL selects the variants without PReLU; nb == 1
 * with weight (Co,Ci,5,5) is the non-batch form.  Bit-exact reduction order. */
int pcx_ctx_conv_step(const float *d_in, const float *d_weight, const float *d_bias, const float *d_act,
                      float *d_out, int nb, int nimg, int npart, int G, int gi, int go, int h, int W,
                      int pad_in, int pad_out, int constrain, int psum, const int *d_order,
                      const int *h_start, void *stream);
/* EntropyAddOp.forward (main.cpp:132-136 -> entropy_add_cuda.cu:47-75), y += x in place. */
int pcx_ctx_add_step(float *d_y, const float *d_x, int nrep, int npart, int G, int cpg, int h, int W, int pad,
                     int psum, const int *d_order, const int *h_start, void *stream);
/* DInput2Op.forward (main.cpp:75-79 -> d_input_cuda_v2.cu:55-86); psum = raw step counter. */
int pcx_dinput_step(const float *d_sym, float *d_out, int nimg, int npart, int G, int h, int W, int pad,
                    float bias, int rep, int psum, const int *d_order, const int *h_start, void *stream);
/* DExtract2Op.forward / forward_batch (main.cpp:68-73 -> d_extract_cuda_v2.cu:54-106, :134-166).
 * lag = 1 reproduces the label==false branch (:85-98).  *h_count = rows produced. */
int pcx_dextract_step(const float *d_in, float *d_out, int nrep, int npart, int G, int cpn, int h, int W,
                      int psum, int batch, int lag, const int *d_order, const int *h_start, int *h_count,
                      void *stream);
/* Fused wavefront step (B200-native): DInput2 + 12x(ctx pad + masked conv) + 5x add + DExtract2Batch +
 * GMM table for one step in ONE launch per layer group; same arithmetic as the separate entry points. */

/* ---- wavefront engine (B200-native): the serial loops of EntEncoder.forward / EntDecoder.forward
 * (pseudo_codec.py:97-114, :145-160) in one native call - same kernels, same CDFs, no per-operator host round trip; int32
 * CDF rows of the live symbols only cross PCIe (pinned, asynchronous) and the host coder runs in the same loop, overlapped
 * with the next step's kernels when encoding. */
#define PCX_WAVE_MAX_LAYERS 16
typedef struct pcx_wave_layer {
    const float *weight;     /* (nb, G*go, G*gi, 5, 5)  EntropyConv2Batch.weight */
    const float *bias;       /* (nb, G*go) */
    const float *act;        /* (nb, G*go) PReLU slopes or NULL */
    float *in;               /* (nb*nimg*npart, G*gi, h+2pad, W+2pad) padded input, halo refreshed in place */
    float *out;              /* (nb*nimg*npart, G*go, h+2pad_out, W+2pad_out) */
    const float *add;        /* EntropyAdd source added to `out` at the wavefront cells (same shape) or NULL */
    int gi, go, pad_out, constrain, input_layer;
} pcx_wave_layer;
typedef struct pcx_wave_net {
    int nlayers, nb, nimg, npart, G, h, W, pad, nstep, ng;
    int cdf_rows;                                    /* capacity of d_cdf / d_lab in rows */
    float gmm_bias, gmm_total, gmm_beta, input_bias;
    const int *wl;                                   /* host, npart */
    const int *d_band, *d_row, *d_col;               /* mode-1 halo table (pcx_halo_table) */
    const float *d_tw;
    const int *d_items;                              /* pcx_ctx_pad_items */
    const int *h_pstart;
    const int *d_order;                              /* pcx_ctx_order */
    const int *h_start;
    float *d_params;                                 /* (nb, go_last, h*npart, W) * nimg extraction buffer */
    int *d_cdf;                                      /* (rows, nstep+1) int32 */
    float *d_prev;                                   /* (nimg, h*npart*W) symbols of the previous step */
    int *d_lab;                                      /* (cdf_rows) int32 symbols in coding order (one-shot encoder) */
    int *d_steptab;                                  /* 2*(steps+1) ints of scratch (one-shot encoder) */
    pcx_wave_layer layers[PCX_WAVE_MAX_LAYERS];
} pcx_wave_net;
int pcx_wave_steps(const pcx_wave_net *net);         /* h*npart + W + G - 2 */
/* d_data (nimg*npart, G, h, W): symbols as float (PseudoFill'ed).  Image i is coded into coders[i] (already started): the
 * nimg images advance through the wavefront together - same step count, nimg times the work per launch - and each
 * gets its own headerless bitstream, identical to coding it alone. */
int pcx_wave_encode(const pcx_wave_net *net, const float *d_data, pcx_coder *const *coders, long long *n_symbols, void *stream);
/* One-shot form of pcx_wave_encode (the encoder knows every symbol): each of the 12 layers is evaluated over the whole tensor
 * in one launch with the per-scalar arithmetic of the stepwise form, the CDF rows are emitted in the stepwise coding order in
 * chunks of at most cdf_rows rows and coded on host threads (one per image) behind the device.  Same bitstreams, byte for byte. */
int pcx_wave_encode_full(const pcx_wave_net *net, const float *d_data, pcx_coder *const *coders, long long *n_symbols, void *stream);
/* Decodes every symbol of the nimg bitstreams; on return layers[0].in holds symbol + input_bias at every valid cell. */
int pcx_wave_decode(const pcx_wave_net *net, pcx_coder *const *coders, long long *n_symbols, void *stream);
/* Engine behind pcx_wave_decode.  2 (default): ONE persistent dataflow kernel per decode - layers synchronise scalar by scalar
 * through write-once scratch, images of a batch are independent pipelines, CDF rows (16 bytes) and symbols cross PCIe through
 * mapped pinned memory, host decoder threads poll them (pcx_flow.cu).  1: one cooperative launch per wavefront step (DInput2,
 * the masked layers with the halo taps interpolated on the fly, residual adds, DExtract2Batch + GMM table, separated by grid
 * barriers).  0: the launch-per-operator sequence.  Same CDFs, symbols and error behaviour in all three.  Returns the previous
 * setting. */
int pcx_wave_set_fused(int mode);
/* Host decoder threads of engine 2 per call (0 = automatic: host cores / LOCAL_WORLD_SIZE - 1, at most one per image; the
 * PCX_CODER_THREADS environment variable overrides the automatic choice).  Returns the previous value. */
int pcx_flow_set_threads(int n);
/* Tuning knobs of the one-shot encoder (pcx_wave_encode_full), for A/B timing and for the tests that pin every variant to the same
 * bytes: "slabs" = number of wavefront slabs the tensor is encoded in (0 = automatic: 1 below 600 steps, 2 below 1400, else 3),
 * "tsplit" = blocks sharing the channel-group pairs of a tile in the shared-memory context convolution (0 = cost model),
 * "smem" = 0 disables that kernel (L1-resident form).  Returns the previous value, or PCX_EINVAL for an unknown name. */
int pcx_wave_set_option(const char *name, int value);

/* ---- GMM ---------------------------------------------------------------------------------------------
 * EntropyGmmTableOp.forward_batch / forward (main.cpp:49-53 -> entropy_gmm_table_cuda.cu:107-185).
 * Softmax and delta clamp are applied IN PLACE on the inputs like the reference; d_cdf_f (n, nstep+1) fp32
 * integer-valued table (reference layout) and/or d_cdf_i int32 (what the coder consumes); either may be NULL. */
/* form 0 = arithmetic of entropy_gmm_table_batch_forward_kernel (:136-153, what the codec uses),
 * form 1 = entropy_gmm_table_forward_kernel (:59-80): the two reference kernels round differently. */
int pcx_gmm_table(float *d_logit, float *d_delta, const float *d_mean, int n, int ng, int nstep, float bias,
                  float total, float beta, int form, float *d_cdf_f, int *d_cdf_i, void *stream);
/* EntropyGmmOp.forward (main.cpp:24-28 -> entropy_gmm_cuda.cu:72-92): loss only. */
int pcx_gmm_nll(const float *d_w, const float *d_delta, const float *d_mean, const float *d_label,
                float *d_loss, int n, int ng, void *stream);

/* ---- evaluation metrics of `pseudo_codec.py --test` --------------------------------
 * ProjectsOp (main.cpp:30-35 -> extension/projects_cuda.cu): 14 rectilinear viewports of an ERP image.
 * pcx_project_table: projects_opt::init + update (:96-152) - theta / phi / fov in units of pi (projects.hpp:8-19);
 * d_tf (14, h_out*w_out, 2) ERP pixel coordinates.  pcx_project_fwd: forward_cuda (:215-246), bilinear (wrap in longitude,
 * clamp in latitude) or nearest; d_out is (14, N*C, h_out, w_out) exactly as the reference kernel writes it. */
int pcx_project_table(const float *theta, const float *phi, float fov, int h_out, int w_out, int H, int W, float *d_tf, void *stream);
int pcx_project_fwd(const float *d_in, const float *d_tf, float *d_out, int N, int C, int H, int W, int h_out, int w_out, int nearest,
                    void *stream);
/* Gaussian-window SSIM (PCONV_operator/pytorch_ssim.py:17-37: window 11, sigma 1.5, zero padding, C1 = 0.01^2, C2 = 0.03^2):
 * *d_mean = mean of the SSIM map over planes*h*w; d_map (optional) receives the map; d_scratch: >= ceil(planes*h*w/256)
 * doubles.  pcx_mean_sqdiff: mean((a-b)^2) (d_b != NULL, the viewport MSE of pseudo_codec.py:276) or mean(a). */
int pcx_ssim(const float *d_a, const float *d_b, long long planes, int h, int w, int window, float sigma, float *d_map,
             double *d_scratch, long long scratch_len, double *d_mean, void *stream);
int pcx_mean_sqdiff(const float *d_a, const float *d_b, long long total, double *d_scratch, long long scratch_len, double *d_mean,
                    void *stream);

/* ---- host arithmetic coder (coder/python.cpp:63-72 `coder.coder`) --------------------------------------
 * 32-bit-state range coder, MSB-first bit stream, no header, one terminating 1 bit then zero padding
 * (coder/ArithmeticCoder.cpp:34-69, :82-116, :152-154; coder/BitIoStream.cpp:52-72). */
pcx_coder *pcx_coder_open(const char *path);          /* coder.coder(path)            */
void pcx_coder_close(pcx_coder *c);
int pcx_coder_start_encoder(pcx_coder *c);             /* start_encoder                */
int pcx_coder_encodes(pcx_coder *c, const int32_t *table, int ncode, const int32_t *symbols, int n); /* encodes */
int pcx_coder_end_encoder(pcx_coder *c);               /* end_encoder                  */
int pcx_coder_start_decoder(pcx_coder *c);             /* start_decoder                */
int pcx_coder_decodes(pcx_coder *c, const int32_t *table, int ncode, int n, float *out_symbols);    /* decodes */
/* coder.decodes (coder/python.cpp:41-60) for the persistent decoder kernel: `rows` holds one 16-byte record per symbol written by
 * the device into mapped pinned memory - cum[1..7] as uint16 (cum[0] = 0, cum[8] = 65536 implied) + a 16-bit tag - and rows
 * [0, n) are decoded for as long as their tag equals tag16 (i.e. as far as the device has got); symbol i goes to out_words[i] as
 * (word_tag << 8 | symbol), the word the device polls.  *done = rows consumed.  Same arithmetic as pcx_coder_decodes. */
int pcx_coder_decodes_rows16(pcx_coder *c, const uint16_t *rows, int n, unsigned tag16, unsigned word_tag, uint32_t *out_words, int *done);
/* in-memory variants for pipelines that keep bitstreams in host RAM */
int pcx_coder_start_encoder_mem(pcx_coder *c);
long long pcx_coder_take_bytes(pcx_coder *c, unsigned char *dst, long long cap);
int pcx_coder_start_decoder_mem(pcx_coder *c, const unsigned char *src, long long n);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* PCX_H */
