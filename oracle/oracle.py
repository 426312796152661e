"""numpy front-end of the CPU oracle (oracle/pcx_oracle.c).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.  Each wrapper names the reference lines restated by
the C function it calls; the arithmetic lives in pcx_oracle.c.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpcx_oracle.so")

# PCONV_operator/base.py:13-35 set_weight(16, opt=True) evaluated with the default 32-entry profile
# (base.py:10); SURVEY.md fact 3.  The product recomputes this through scipy; the oracle pins it.
W64_NPART16 = [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]


def build(force=False):
    src = os.path.join(_HERE, "pcx_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libpcx_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(C.POINTER(C.c_int))


# ---------------------------------------------------------------------------------------- geometry
def band_widths(weight, H, W):
    """extension/math_cuda.cu:223-253."""
    w, wp = _f(weight)
    out = np.zeros(len(w), np.int32)
    rc = lib().orc_band_widths(wp, len(w), int(H), int(W), _ip(out))
    if rc != 0:
        raise ValueError("height must be a multiple of npart (math_cuda.cu:225)")
    return out


def slice_table(wl, W):
    wl, wp = _i(wl)
    src = np.zeros((len(wl), W), np.int32)
    wt = np.zeros((len(wl), W, 4), np.float32)
    lib().orc_slice_table(wp, len(wl), int(W), _ip(src), _fp(wt))
    return src, wt


def uslice_table(wl, W):
    wl, wp = _i(wl)
    src = np.zeros((len(wl), W), np.int32)
    wt = np.zeros((len(wl), W, 4), np.float32)
    lib().orc_uslice_table(wp, len(wl), int(W), _ip(src), _fp(wt))
    return src, wt


def sphere_slice(x, wl, pad=0):
    """sphere_slice_cuda.cu:87-116.  x (N,C,H,W) -> (N*npart, C, H/npart+2pad, W+2pad)."""
    x, xp = _f(x)
    wl, wp = _i(wl)
    N, Cc, H, W = x.shape
    npart = len(wl)
    src, wt = slice_table(wl, W)
    out = np.zeros((N * npart, Cc, H // npart + 2 * pad, W + 2 * pad), np.float32)
    lib().orc_slice(xp, _fp(out), N, Cc, H, W, npart, wp, _ip(src), _fp(wt), int(pad))
    return out


def sphere_uslice(x, wl, pad=0):
    """sphere_uslice_cuda.cu:73-99.  x (N*npart, C, h+2pad, W+2pad) -> (N, C, h*npart, W)."""
    x, xp = _f(x)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, hh, ww = x.shape
    h, W = hh - 2 * pad, ww - 2 * pad
    N = NN // npart
    src, wt = uslice_table(wl, W)
    out = np.zeros((N, Cc, h * npart, W), np.float32)
    lib().orc_uslice(xp, _fp(out), N, Cc, h, W, npart, wp, _ip(src), _fp(wt), int(pad))
    return out


def halo_table(wl, h, W, pad, mode=0):
    """mode 0: pseudo_context_cuda.cu:51-104; 1: entropy_context_cuda.cu:105-165 (== pseudo_entropy v1);
    2: pseudo_entropy_context_cuda.cu:50-109 (v0)."""
    wl, wp = _i(wl)
    npart = len(wl)
    band = np.zeros((npart, 2, pad), np.int32)
    row = np.zeros((npart, 2, pad), np.int32)
    col = np.zeros((npart, 2, pad, W), np.int32)
    tw = np.zeros((npart, 2, pad, W), np.float32)
    lib().orc_halo_table(wp, npart, int(h), int(W), int(pad), int(mode), _ip(band), _ip(row), _ip(col), _fp(tw))
    return band, row, col, tw


def pseudo_pad(x, wl, pad):
    """pseudo_pad.cu:39-96."""
    x, xp = _f(x)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, h, W = x.shape
    band, row, col, tw = halo_table(wl, h, W, pad, 0)
    out = np.zeros((NN, Cc, h + 2 * pad, W + 2 * pad), np.float32)
    lib().orc_pad(xp, _fp(out), NN // npart, Cc, h, W, npart, int(pad), wp, _ip(band), _ip(row), _ip(col), _fp(tw))
    return out


def pseudo_entropy_pad(x, wl, pad, version=1):
    """pseudo_entropy_pad_cuda.cu:39-105 with the v1 (default) or v0 causal table."""
    x, xp = _f(x)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, h, W = x.shape
    band, row, col, tw = halo_table(wl, h, W, pad, 1 if version == 1 else 2)
    out = np.zeros((NN, Cc, h + 2 * pad, W + 2 * pad), np.float32)
    lib().orc_entropy_pad(xp, _fp(out), NN // npart, Cc, h, W, npart, int(pad), wp, _ip(band), _ip(row), _ip(col), _fp(tw))
    return out


def pseudo_fill(x, wl, pad=0, trim=0, fvalue=0.0):
    """pseudo_fill_cuda.cu:28-43 (returns a filled copy)."""
    x = np.array(x, dtype=np.float32, order="C", copy=True)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, Hh, Ww = x.shape
    lib().orc_fill(_fp(x), NN // npart, Cc, Hh, Ww, npart, int(pad), int(trim), wp, C.c_float(fvalue))
    return x


# ---------------------------------------------------------------------------------------- quantiser
def quant_steps(theta):
    theta, tp = _f(theta)
    w = np.zeros_like(theta)
    lib().orc_quant_steps(tp, _fp(w), theta.shape[0], theta.shape[1])
    return w


def dquant_centres(theta):
    theta, tp = _f(theta)
    c = np.zeros_like(theta)
    lib().orc_dquant_centres(tp, _fp(c), theta.shape[0], theta.shape[1])
    return c


def pseudo_quant(x, steps, wl):
    """pseudo_quant_cuda.cu:48-94.  `steps` = expanded table (quant_steps or the GPU's own)."""
    x, xp = _f(x)
    steps, sp = _f(steps)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, h, W = x.shape
    val = np.zeros_like(x)
    sym = np.zeros_like(x)
    count = np.zeros_like(steps)
    lib().orc_quant(xp, _fp(val), _fp(sym), _fp(count), sp, NN // npart, Cc, h, W, npart, steps.shape[1], wp)
    return val, sym, count


def pseudo_dquant(sym, centres, wl):
    """pseudo_dquant_cuda.cu:34-47."""
    sym, sp = _f(sym)
    centres, cp = _f(centres)
    wl, wp = _i(wl)
    npart = len(wl)
    NN, Cc, h, W = sym.shape
    out = np.zeros_like(sym)
    lib().orc_dquant(sp, _fp(out), cp, NN // npart, Cc, h, W, npart, centres.shape[1], wp)
    return out


def dtow(x, stride=2, d2w=True):
    """dtow_cuda.cu:38-75."""
    x, xp = _f(x)
    N, Cc, H, W = x.shape
    s = stride
    if d2w:
        out = np.zeros((N, Cc // (s * s), H * s, W * s), np.float32)
    else:
        out = np.zeros((N, Cc * s * s, H // s, W // s), np.float32)
    lib().orc_dtow(xp, _fp(out), N, Cc, H, W, s, 1 if d2w else 0)
    return out


def conv2d(x, w, b=None, stride=1):
    """fp64-accumulated direct convolution (tolerance anchor; the reference calls cuDNN)."""
    x, xp = _f(x)
    w, wp = _f(w)
    N, Ci, Hi, Wi = x.shape
    Co, _, k, _ = w.shape
    Ho, Wo = (Hi - k) // stride + 1, (Wi - k) // stride + 1
    y = np.zeros((N, Co, Ho, Wo), np.float32)
    bp = None
    if b is not None:
        b, bp = _f(b)
    lib().orc_conv2d(xp, wp, bp, _fp(y), N, Ci, Hi, Wi, Co, k, int(stride))
    return y


def gdn(x, beta, gamma, wl, inverse=False, beta_min=1e-6, reparam_offset=2.0 ** -18):
    """PCONV_operator/PseudoContextV2.py:186-216 + GDN.py:6-22 (LowerBound), float64 accumulation."""
    x = np.asarray(x, np.float32)
    npart = len(wl)
    NN, Cc, h, W = x.shape
    pedestal = np.float32(reparam_offset) ** 2
    beta_bound = np.float32((beta_min + reparam_offset ** 2) ** 0.5)
    gamma_bound = np.float32(reparam_offset)
    b = np.maximum(np.asarray(beta, np.float32), beta_bound) ** 2 - pedestal
    g = np.maximum(np.asarray(gamma, np.float32), gamma_bound) ** 2 - pedestal
    mask = np.zeros((NN, 1, 1, W), np.float32)
    for n in range(NN):
        mask[n, 0, 0, : wl[n % npart]] = 1
    xm = x * mask
    norm = np.einsum("oc,nchw->nohw", g.astype(np.float64), (xm.astype(np.float64)) ** 2) + b.astype(np.float64)[None, :, None, None]
    norm = np.sqrt(norm).astype(np.float32)
    norm = norm * mask + 1 - mask
    return (xm * norm if inverse else xm / norm).astype(np.float32)


# --------------------------------------------------------------------------------- context model
class CtxGeom:
    """entropy_context (entropy_context.hpp:10-50, entropy_context_cuda.cu:13-45, :168-221) for one
    code width: wavefront order, causal halo table and the per-plane pad work lists."""

    def __init__(self, wl, h, W, pad=2):
        self.wl, _ = _i(wl)
        self.npart = len(self.wl)
        self.h, self.W, self.pad = int(h), int(W), int(pad)
        self.Hf = self.h * self.npart
        self.order = np.zeros(self.Hf * self.W, np.int32)
        self.start = np.zeros(self.Hf + self.W, np.int32)
        lib().orc_ctx_order(_ip(self.wl), self.npart, self.h, self.W, _ip(self.order), _ip(self.start))
        self.band, self.row, self.col, self.tw = halo_table(self.wl, h, W, pad, 1)
        self.pstart = np.zeros(self.Hf + self.W + self.pad, np.int32)
        n = lib().orc_ctx_pad_items(_ip(self.wl), self.npart, self.h, self.W, self.pad, _ip(self.band),
                                    _ip(self.col), _fp(self.tw), None, _ip(self.pstart))
        self.items = np.zeros((max(n, 1), 4), np.int32)
        lib().orc_ctx_pad_items(_ip(self.wl), self.npart, self.h, self.W, self.pad, _ip(self.band),
                                _ip(self.col), _fp(self.tw), _ip(self.items), _ip(self.pstart))

    def nsteps(self, G):
        return self.Hf + self.W + G - 2

    def window(self, psum, G):
        st = max(0, psum - G + 1)
        en = psum + 1 if psum < self.Hf + self.W - 2 else self.Hf + self.W - 1
        return int(self.start[st]), int(self.start[en])


def ctx_pad_step(buf, geom, G, psum):
    """entropy_ctx_pad_run2_cuda.cu:33-65, :86-117 (in place; caller applies the input-layer lag)."""
    assert buf.dtype == np.float32 and buf.flags.c_contiguous
    NN, Cc, hh, ww = buf.shape
    lib().orc_ctx_pad_step(_fp(buf), NN // geom.npart, geom.npart, int(G), Cc // G, geom.h, geom.W, geom.pad,
                           int(psum), _ip(geom.wl), _ip(geom.band), _ip(geom.row), _ip(geom.col), _fp(geom.tw),
                           _ip(geom.items), _ip(geom.pstart))
    return buf


def ctx_conv_step(x, weight, bias, act, out, geom, G, nimg, pad_in, pad_out, constrain, psum, pool=None, threads=1):
    """entropy_conv_cuda_v2.cu:237-290 / :326-379 with the A.6 reduction tree."""
    assert x.dtype == np.float32 and out.dtype == np.float32
    weight, wp = _f(weight)
    bias, bp = _f(bias)
    ap = None
    if act is not None:
        act, ap = _f(act)
    nb = weight.shape[0]
    go = weight.shape[1] // G
    gi = weight.shape[2] // G
    args = (_fp(x), wp, bp, ap, _fp(out), nb, int(nimg), geom.npart, int(G), gi, go, geom.h, geom.W,
            int(pad_in), int(pad_out), int(constrain), int(psum), _ip(geom.order), _ip(geom.start))
    lo, hi = geom.window(psum, G)
    ntask = nb * int(nimg) * (hi - lo)
    if pool is not None and threads > 1 and ntask >= 4 * threads and psum < geom.nsteps(G):
        # CPU baseline: the step's independent scalars spread over the host cores (ctypes releases the GIL)
        step = -(-ntask // threads)
        list(pool.map(lambda a: lib().orc_ctx_conv_step_range(*args, C.c_int64(a), C.c_int64(min(a + step, ntask))),
                      range(0, ntask, step)))
    else:
        lib().orc_ctx_conv_step(*args)
    return out


def ctx_add_step(y, x, geom, G, pad, psum):
    """entropy_add_cuda.cu:25-44."""
    NN, Cc, _, _ = y.shape
    lib().orc_ctx_add_step(_fp(y), _fp(x), NN // geom.npart, geom.npart, int(G), Cc // G, geom.h, geom.W, int(pad),
                           int(psum), _ip(geom.order), _ip(geom.start))
    return y


def dinput_step(sym, out, geom, G, nimg, pad, bias, rep, psum):
    """d_input_cuda_v2.cu:32-52, :55-86."""
    sym, sp = _f(sym)
    lib().orc_dinput_step(sp, _fp(out), int(nimg), geom.npart, int(G), geom.h, geom.W, int(pad), C.c_float(bias),
                          int(rep), int(psum), _ip(geom.order), _ip(geom.start))
    return out


def dextract_step(x, out, geom, G, psum, batch):
    """d_extract_cuda_v2.cu:34-52 / :110-132.  Returns the symbol count the op reports."""
    NN, Cc, _, _ = x.shape
    return lib().orc_dextract_step(_fp(x), _fp(out), NN // geom.npart, geom.npart, int(G), Cc // G, geom.h, geom.W,
                                   int(psum), 1 if batch else 0, _ip(geom.order), _ip(geom.start))


def gmm_table(logit, delta, mean, nstep=8, bias=3.5, total=65536.0, beta=1e-6, form=0):
    """entropy_gmm_table_cuda.cu:29-56, :83-105, :136-153 -> int32 (n, nstep+1)."""
    logit, lp = _f(logit)
    delta, dp = _f(delta)
    mean, mp = _f(mean)
    n, ng = logit.shape
    cdf = np.zeros((n, nstep + 1), np.int32)
    w = np.zeros_like(logit)
    d = np.zeros_like(logit)
    lib().orc_gmm_table(lp, dp, mp, n, ng, int(nstep), C.c_float(bias), C.c_float(total), C.c_float(beta), int(form),
                        _ip(cdf), _fp(w), _fp(d))
    return cdf, w, d


def gmm_nll(w, delta, mean, label):
    """entropy_gmm_cuda.cu:36-69 (forward value)."""
    w, wp = _f(w)
    delta, dp = _f(delta)
    mean, mp = _f(mean)
    label, lp = _f(label)
    n, ng = w.shape
    loss = np.zeros(n, np.float32)
    lib().orc_gmm_nll(wp, dp, mp, lp, _fp(loss), n, ng)
    return loss
