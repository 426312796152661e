"""ctypes binding of libpcx.so - the C ABI declared in include/pcx.h.

There is no fallback: if the library is missing or a call fails, PcxError is raised.  The library is built
in-tree by `pseudocylindrical_convolution_b200.build` (nvcc, sm_100a).
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
# PCX_LIB: an alternative build of the same ABI (A/B timing of kernel variants); the default is the in-tree library
LIB_PATH = os.environ.get("PCX_LIB") or os.path.join(_PKG, "libpcx.so")

PCX_MAX_PART = 32


class PcxError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    """struct pcx_conv_desc (include/pcx.h)."""
    _fields_ = [(n, C.c_int) for n in (
        "N", "npart", "Ci", "Hi", "in_pitch", "Co", "Ho", "Wo", "out_rows", "out_pitch", "out_y0", "out_x0",
        "k", "stride", "act", "impl", "aux_rows", "aux_pitch", "aux_y0", "aux_x0", "in_plane_rows")] + [("wl_out", C.c_int * PCX_MAX_PART),
                                                                                                          ("square_input", C.c_int)]


class WaveLayer(C.Structure):
    """struct pcx_wave_layer (include/pcx.h)."""
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("act", C.c_void_p), ("in_", C.c_void_p), ("out", C.c_void_p),
                ("add", C.c_void_p)] + [(n, C.c_int) for n in ("gi", "go", "pad_out", "constrain", "input_layer")]


class WaveNet(C.Structure):
    """struct pcx_wave_net (include/pcx.h)."""
    _fields_ = ([(n, C.c_int) for n in ("nlayers", "nb", "nimg", "npart", "G", "h", "W", "pad", "nstep", "ng", "cdf_rows")] +
                [(n, C.c_float) for n in ("gmm_bias", "gmm_total", "gmm_beta", "input_bias")] +
                [(n, C.c_void_p) for n in ("wl", "d_band", "d_row", "d_col", "d_tw", "d_items", "h_pstart", "d_order", "h_start",
                                           "d_params", "d_cdf", "d_prev", "d_lab", "d_steptab")] +
                [("layers", WaveLayer * 16)])


_P = C.c_void_p
_I = C.c_int
_F = C.c_float
_IP = C.POINTER(C.c_int)
_FP = C.POINTER(C.c_float)

# name -> (restype, argtypes).  Every symbol declared in include/pcx.h must appear here
# (tests/test_abi.py checks the header against this table and against the built library).
PROTOTYPES = {
    "pcx_abi_version": (_I, []),
    "pcx_last_error": (C.c_char_p, []),
    "pcx_launch_count": (C.c_longlong, []),
    "pcx_device_check": (_I, [_I, _IP, _IP]),
    "pcx_band_widths": (_I, [_FP, _I, _I, _I, _IP]),
    "pcx_slice_table": (_I, [_IP, _I, _I, _P, _P, _P]),
    "pcx_uslice_table": (_I, [_IP, _I, _I, _P, _P, _P]),
    "pcx_halo_table": (_I, [_IP, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "pcx_slice_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _IP, _P, _P, _I, _P]),
    "pcx_uslice_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _IP, _P, _P, _I, _P]),
    "pcx_pad_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _I, _P]),
    "pcx_entropy_pad_fwd": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _P]),
    "pcx_halo_fill": (_I, [_P, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _I, _P]),
    "pcx_slice_pad_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "pcx_uslice_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P]),
    "pcx_halo_fill_nhwc": (_I, [_P, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _I, _I, _I, _I, _P]),
    "pcx_dtow_nhwc": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pcx_square": (_I, [_P, _P, C.c_longlong, _P]),
    "pcx_fill": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _IP, _F, _P]),
    "pcx_dtow": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "pcx_quant_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _IP, _P]),
    "pcx_dquant_fwd": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _IP, _P]),
    "pcx_conv2d_fwd": (_I, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "pcx_conv_pack_weights": (C.c_longlong, [_P, _P, _I, _I, _I, _P]),
    "pcx_conv_pack_weights_d2w": (C.c_longlong, [_P, _P, _I, _I, _I, _P]),
    "pcx_gdn_params": (_I, [_P, _P, _P, _P, _I, _F, _F, _P]),
    "pcx_gdn_fwd": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _IP, _I, _P]),
    "pcx_ctx_order": (_I, [_IP, _I, _I, _I, _IP, _IP]),
    "pcx_ctx_pad_items": (_I, [_IP, _I, _I, _I, _I, _IP, _IP, _FP, _IP, _IP]),
    "pcx_ctx_pad_step": (_I, [_P, _I, _I, _I, _I, _I, _I, _I, _I, _IP, _P, _P, _P, _P, _P, _IP, _P]),
    "pcx_ctx_conv_step": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _IP, _P]),
    "pcx_ctx_add_step": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _P, _IP, _P]),
    "pcx_dinput_step": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _F, _I, _I, _P, _IP, _P]),
    "pcx_dextract_step": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P, _IP, _IP, _P]),
    "pcx_wave_steps": (_I, [C.POINTER(WaveNet)]),
    "pcx_wave_encode": (_I, [C.POINTER(WaveNet), _P, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), _P]),
    "pcx_wave_encode_full": (_I, [C.POINTER(WaveNet), _P, C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), _P]),
    "pcx_wave_set_fused": (_I, [_I]),
    "pcx_flow_set_threads": (_I, [_I]),
    "pcx_wave_set_option": (_I, [C.c_char_p, _I]),
    "pcx_wave_decode": (_I, [C.POINTER(WaveNet), C.POINTER(C.c_void_p), C.POINTER(C.c_longlong), _P]),
    "pcx_gmm_table": (_I, [_P, _P, _P, _I, _I, _I, _F, _F, _F, _I, _P, _P, _P]),
    "pcx_gmm_nll": (_I, [_P, _P, _P, _P, _P, _I, _I, _P]),
    "pcx_project_table": (_I, [_FP, _FP, _F, _I, _I, _I, _I, _P, _P]),
    "pcx_project_fwd": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
    "pcx_ssim": (_I, [_P, _P, C.c_longlong, _I, _I, _I, _F, _P, _P, C.c_longlong, _P, _P]),
    "pcx_mean_sqdiff": (_I, [_P, _P, C.c_longlong, _P, C.c_longlong, _P, _P]),
    "pcx_coder_open": (_P, [C.c_char_p]),
    "pcx_coder_close": (None, [_P]),
    "pcx_coder_start_encoder": (_I, [_P]),
    "pcx_coder_encodes": (_I, [_P, _P, _I, _P, _I]),
    "pcx_coder_end_encoder": (_I, [_P]),
    "pcx_coder_start_decoder": (_I, [_P]),
    "pcx_coder_decodes": (_I, [_P, _P, _I, _I, _P]),
    "pcx_coder_decodes_rows16": (_I, [_P, _P, _I, C.c_uint, C.c_uint, _P, _IP]),
    "pcx_coder_start_encoder_mem": (_I, [_P]),
    "pcx_coder_take_bytes": (C.c_longlong, [_P, _P, C.c_longlong]),
    "pcx_coder_start_decoder_mem": (_I, [_P, _P, C.c_longlong]),
}

_lib = None


def load():
    """Load libpcx.so (once).  Raises PcxError if it has not been built - there is no software fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PcxError(
            "libpcx.so is missing (%s). Build it with `python -m pseudocylindrical_convolution_b200.build`; "
            "this package has no CPU or PyTorch fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Turn a negative status into PcxError carrying pcx_last_error()."""
    if rc is not None and rc < 0:
        raise PcxError("libpcx error %d: %s" % (rc, load().pcx_last_error().decode("utf-8", "replace")))
    return rc


def call(name, *args):
    return check(getattr(load(), name)(*args))


def int_array(values):
    arr = (C.c_int * len(values))(*[int(v) for v in values])
    return arr
