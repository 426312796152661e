"""`PCONV` - host-side mirror of the reference's pybind11 extension module (extension/main.cpp:4-137).

Same class names, positional constructor signatures, method names, return conventions and ownership rules
as the reference classes, so `PCONV_operator/*.py` (the reference's or this package's) runs on top of it
unchanged.  The compute goes through the C ABI of libpcx.so (include/pcx.h); this file owns what the
reference objects own on the host: cached output buffers (extension/base_opt.hpp:43-72), the wavefront step
counters `pidx_` with restart() (e.g. extension/entropy_conv_v2.hpp:27), and the per-shape table caches of
the three context objects (extension/pseudo_context.hpp:8-43, entropy_context.hpp:10-50).

Differences, all deliberate:
  * context "addresses" are registry keys, not raw pointers (string2class.cc:2-22 reinterpret_casts a hex
    string); dependants keep a strong reference, so the shell may be collected first without a dangling pointer;
  * only float32 CUDA tensors are accepted (the reference also dispatches double);
  * backward() methods raise: training is out of scope;
  * errors raise PcxError instead of printf (extension/caffe_cuda_macro.h:21-26).
"""
import ctypes as C
import itertools

import numpy as np
import torch

from . import _lib
from ._lib import PcxError, call, int_array

__all__ = [
    "SphereSliceOp", "SphereUsliceOp", "PseudoContextOp", "PseudoPadOp", "PseudoFillOp", "PseudoQuantOp",
    "PseudoDQuantOp", "DtowOp", "EntropyContextOp", "EntropyCtxPadRun2Op", "EntropyConv2Op", "EntropyAddOp",
    "DInput2Op", "DExtract2Op", "EntropyGmmTableOp", "EntropyGmmOp", "PseudoEntropyContextOp",
    "PseudoEntropyPadOp", "ProjectsOp", "ContextReshapeOp", "MaskConstrainOp",
]


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check_tensor(t, name="input"):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise TypeError("%s must be a float32 CUDA tensor" % name)
    if not t.is_contiguous():
        raise ValueError("%s must be contiguous" % name)
    return t


class _Op:
    """base_opt (extension/base_opt.hpp): device binding + output buffers reallocated only on shape change."""

    def __init__(self, device, timeit=False):
        self.device_ = int(device)
        self.timeit_ = bool(timeit)
        self._outs = {}
        _lib.load()

    def to(self, device):
        if int(device) != self.device_:
            self.device_ = int(device)
            self._outs.clear()
            self._on_device_change()

    def _on_device_change(self):
        pass

    def _out(self, key, shape, like, dtype=None, zero=False):
        shape = tuple(int(s) for s in shape)
        buf = self._outs.get(key)
        if buf is None or tuple(buf.shape) != shape or buf.device != like.device:
            buf = (torch.zeros if zero else torch.empty)(shape, dtype=dtype or like.dtype, device=like.device)
            self._outs[key] = buf
        return buf

    def backward(self, *a, **k):
        raise NotImplementedError("%s.backward: the training path is out of scope of this implementation" % type(self).__name__)


# ------------------------------------------------------------------------------------------------ contexts
_registry = {}
_addr_counter = itertools.count(1)


def _resolve(addr):
    try:
        return _registry[addr]
    except KeyError:
        raise PcxError("unknown context address %r (contexts are process-local, like the reference's pointers)" % (addr,))


class _Geometry:
    """Band widths and gather tables shared by address, cached per width like the reference's std::map caches."""

    halo_mode = 0

    def __init__(self, npart, rt, weight, device, timeit):
        self.npart_ = int(npart)
        self.rt_ = int(rt)
        self.weight_ = [float(w) for w in weight][: self.npart_]
        if len(self.weight_) != self.npart_:
            raise ValueError("weight must have npart entries")
        self.device_ = int(device)
        self.data_width_ = -1
        self._wl = {}
        self._halo = {}
        self.addr_ = "pcx-ctx-%d" % next(_addr_counter)
        _registry[self.addr_] = self
        _lib.load()

    # -- mirror of the shell classes
    def to(self, device):
        if int(device) != self.device_:
            self.device_ = int(device)
            self._clear()

    def start_context(self, width):
        if int(width) != self.data_width_:
            self._clear()
        self.data_width_ = int(width)

    def addr(self):
        return self.addr_

    def _clear(self):
        self._wl.clear()
        self._halo.clear()

    def _dev(self):
        return torch.device("cuda", self.device_)

    # -- geometry
    def widths(self, h, W):
        """sphere_cal_npart_hw_v3 (extension/math_cuda.cu:223-253) for tiles of height h and width W."""
        key = int(W)
        if key not in self._wl:
            w = (C.c_float * self.npart_)(*self.weight_)
            out = (C.c_int * self.npart_)()
            call("pcx_band_widths", w, self.npart_, int(h) * self.npart_, int(W), out)
            self._wl[key] = [int(v) for v in out]
        return self._wl[key]

    def produce_fill_param(self, h, W):
        wl = self.widths(h, W)
        return torch.tensor(wl, dtype=torch.int32, device=self._dev())

    def halo(self, channel, h, W, pad):
        """Halo gather table for (W, pad) - produce_param (extension/pseudo_context_cuda.cu:140-161).  The
        reference's cache key also holds the channel count because its table stores channel-dependent element
        offsets; ours stores (band, row, column, weight) and is channel-free."""
        if not (channel < 1000 and pad < 10):
            raise PcxError("the channel number should be less than 1000 and the pad size should be less than 10 "
                           "(extension/pseudo_context_cuda.cu:38)")
        key = (int(h), int(W), int(pad))
        if key not in self._halo:
            wl = self.widths(h, W)
            dev = self._dev()
            band = torch.zeros((self.npart_, 2, pad), dtype=torch.int32, device=dev)
            row = torch.zeros((self.npart_, 2, pad), dtype=torch.int32, device=dev)
            col = torch.zeros((self.npart_, 2, pad, W), dtype=torch.int32, device=dev)
            tw = torch.zeros((self.npart_, 2, pad, W), dtype=torch.float32, device=dev)
            with torch.cuda.device(dev):
                call("pcx_halo_table", int_array(wl), self.npart_, int(h), int(W), int(pad), self.halo_mode,
                     _p(band), _p(row), _p(col), _p(tw), _stream())
            self._halo[key] = (band, row, col, tw)
        return self._halo[key]


class PseudoContextOp(_Geometry):
    """pseudo_context_shell (extension/pseudo_context.hpp:46-69): PseudoContextOp(npart, rt, weight, device, timeit)."""
    halo_mode = 0


class PseudoEntropyContextOp(_Geometry):
    """pseudo_entropy_context_shell: PseudoEntropyContextOp(npart, rt, context_version, weight, device, timeit)."""

    def __init__(self, npart, rt, context_version, weight, device=0, timeit=False):
        super().__init__(npart, rt, weight, device, timeit)
        if context_version not in (0, 1):
            raise PcxError("undefined context version (extension/pseudo_entropy_context_cuda.cu:229)")
        self.halo_mode = 1 if context_version == 1 else 2


class EntropyContextOp(_Geometry):
    """entropy_context_shell (extension/entropy_context.hpp:52-74): wavefront order + causal halo work lists."""
    halo_mode = 1

    def __init__(self, npart, rt, weight, device=0, timeit=False):
        super().__init__(npart, rt, weight, device, timeit)
        self._order = {}
        self._items = {}

    def _clear(self):
        super()._clear()
        self._order.clear()
        self._items.clear()

    def order(self, h, W):
        """produce_param_group (extension/entropy_context_cuda.cu:211-215): (device order, host start prefix)."""
        key = (int(h), int(W))
        if key not in self._order:
            wl = self.widths(h, W)
            Hf = h * self.npart_
            order = np.zeros(Hf * W, np.int32)
            start = np.zeros(Hf + W, np.int32)
            call("pcx_ctx_order", int_array(wl), self.npart_, int(h), int(W),
                 order.ctypes.data_as(C.POINTER(C.c_int)), start.ctypes.data_as(C.POINTER(C.c_int)))
            d_order = torch.from_numpy(order).to(self._dev())
            self._order[key] = (d_order, start)
        return self._order[key]

    def pad_items(self, h, W, pad):
        """produce_param (extension/entropy_context_cuda.cu:168-209): per-plane halo / right-wrap work lists."""
        key = (int(h), int(W), int(pad))
        if key not in self._items:
            wl = self.widths(h, W)
            band, row, col, tw = self.halo(1, h, W, pad)
            hb = band.cpu().numpy().reshape(-1).copy()
            hc = col.cpu().numpy().reshape(-1).copy()
            ht = tw.cpu().numpy().reshape(-1).copy()
            Hf = h * self.npart_
            pstart = np.zeros(Hf + W + pad, np.int32)
            ip = C.POINTER(C.c_int)
            n = call("pcx_ctx_pad_items", int_array(wl), self.npart_, int(h), int(W), int(pad), hb.ctypes.data_as(ip),
                     hc.ctypes.data_as(ip), ht.ctypes.data_as(C.POINTER(C.c_float)), None, pstart.ctypes.data_as(ip))
            items = np.zeros((max(n, 1), 4), np.int32)
            call("pcx_ctx_pad_items", int_array(wl), self.npart_, int(h), int(W), int(pad), hb.ctypes.data_as(ip),
                 hc.ctypes.data_as(ip), ht.ctypes.data_as(C.POINTER(C.c_float)), items.ctypes.data_as(ip),
                 pstart.ctypes.data_as(ip))
            self._items[key] = (torch.from_numpy(items).to(self._dev()), pstart)
        return self._items[key]


# ------------------------------------------------------------------------------------------------ tile pipeline
class _Resample(_Op):
    def __init__(self, npart, interp_type, pad, weight, device=0, timeit=False):
        super().__init__(device, timeit)
        self.npart_, self.interp_type_, self.pad_ = int(npart), int(interp_type), int(pad)
        self.weight_ = [float(w) for w in weight][: self.npart_]
        self._tables = {}

    def _on_device_change(self):
        self._tables.clear()

    def _geometry(self, H, W, like, fn):
        key = (int(H), int(W))
        if key not in self._tables:
            w = (C.c_float * self.npart_)(*self.weight_)
            out = (C.c_int * self.npart_)()
            call("pcx_band_widths", w, self.npart_, int(H), int(W), out)
            wl = [int(v) for v in out]
            src = torch.zeros((self.npart_, W), dtype=torch.int32, device=like.device)
            wt = torch.zeros((self.npart_, W, 4), dtype=torch.float32, device=like.device)
            call(fn, int_array(wl), self.npart_, int(W), _p(src), _p(wt), _stream())
            self._tables[key] = (wl, src, wt)
        return self._tables[key]


class SphereSliceOp(_Resample):
    """sphere_slice_opt (extension/sphere_slice.hpp): SphereSliceOp(npart, interp_type, pad, weight, device, timeit)."""

    def forward(self, x):
        _check_tensor(x)
        N, Cc, H, W = x.shape
        if H % self.npart_ != 0:
            raise PcxError("height should be multipler of the number of parts (extension/math_cuda.cu:179)")
        with torch.cuda.device(x.device):
            wl, src, wt = self._geometry(H, W, x, "pcx_slice_table")
            h = H // self.npart_
            out = self._out("top", (N * self.npart_, Cc, h + 2 * self.pad_, W + 2 * self.pad_), x)
            call("pcx_slice_fwd", _p(x), _p(out), N, Cc, H, W, self.npart_, int_array(wl), _p(src), _p(wt), self.pad_, _stream())
        return [out]


class SphereUsliceOp(_Resample):
    """sphere_uslice_opt (extension/sphere_uslice.hpp): SphereUsliceOp(npart, interp_type, pad, weight, device, timeit)."""

    def forward(self, x):
        _check_tensor(x)
        NN, Cc, hh, ww = x.shape
        h, W = hh - 2 * self.pad_, ww - 2 * self.pad_
        if NN % self.npart_ != 0:
            raise PcxError("batch %d is not a multiple of npart %d" % (NN, self.npart_))
        N = NN // self.npart_
        with torch.cuda.device(x.device):
            wl, src, wt = self._geometry(h * self.npart_, W, x, "pcx_uslice_table")
            out = self._out("top", (N, Cc, h * self.npart_, W), x)
            call("pcx_uslice_fwd", _p(x), _p(out), N, Cc, h, W, self.npart_, int_array(wl), _p(src), _p(wt), self.pad_, _stream())
        return [out]


class PseudoPadOp(_Op):
    """pseudo_pad_opt (extension/pseudo_pad.hpp): PseudoPadOp(pad, npart, ctx_addr, device, timeit)."""
    entry = "pcx_pad_fwd"

    def __init__(self, pad, npart, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.pad_, self.npart_ = int(pad), int(npart)
        self.ctx_ = _resolve(ctx_addr)

    def forward(self, x):
        _check_tensor(x)
        NN, Cc, h, W = x.shape
        with torch.cuda.device(x.device):
            wl = self.ctx_.widths(h, W)
            band, row, col, tw = self.ctx_.halo(Cc, h, W, self.pad_)
            out = self._out("top", (NN, Cc, h + 2 * self.pad_, W + 2 * self.pad_), x)
            args = [_p(x), _p(out), NN // self.npart_, Cc, h, W, self.npart_, self.pad_, int_array(wl),
                    _p(band), _p(row), _p(col), _p(tw)]
            if self.entry == "pcx_pad_fwd":
                args.append(W + 2 * self.pad_)
            call(self.entry, *args, _stream())
        return [out]


class PseudoEntropyPadOp(PseudoPadOp):
    """pseudo_entropy_pad_opt: PseudoEntropyPadOp(pad, npart, ctx_addr, device, timeit)."""
    entry = "pcx_entropy_pad_fwd"


class PseudoFillOp(_Op):
    """pseudo_fill_opt (extension/pseudo_fill.hpp): PseudoFillOp(pad, npart, fvalue, trim, addr, context_version, device, timeit)."""

    def __init__(self, pad, npart, fvalue, trim, addr, context_version, device=0, timeit=False):
        super().__init__(device, timeit)
        self.pad_, self.npart_, self.fvalue_, self.trim_ = int(pad), int(npart), int(fvalue), int(trim)
        self.context_version_ = int(context_version)
        self.ctx_ = _resolve(addr)

    def forward(self, x):
        _check_tensor(x)
        NN, Cc, Hh, Ww = x.shape
        with torch.cuda.device(x.device):
            # the reference asks the context for widths with the tensor's own (padded) extent (pseudo_fill_cuda.cu:12-25)
            wl = self.ctx_.widths(Hh, Ww)
            call("pcx_fill", _p(x), NN // self.npart_, Cc, Hh, Ww, self.npart_, self.pad_, self.trim_, int_array(wl),
                 C.c_float(self.fvalue_), _stream())
        return [x]

    def backward(self, g):
        _check_tensor(g)
        NN, Cc, Hh, Ww = g.shape
        with torch.cuda.device(g.device):
            wl = self.ctx_.widths(Hh, Ww)
            call("pcx_fill", _p(g), NN // self.npart_, Cc, Hh, Ww, self.npart_, self.pad_, self.trim_, int_array(wl),
                 C.c_float(0.0), _stream())
        return [g]


class DtowOp(_Op):
    """dtow_opt (extension/dtow.hpp): DtowOp(stride, d2w, device, timeit)."""

    def __init__(self, stride=2, d2w=True, device=0, timeit=False):
        super().__init__(device, timeit)
        self.stride_, self.d2w_ = int(stride), bool(d2w)

    def forward(self, x):
        _check_tensor(x)
        N, Cc, H, W = x.shape
        s = self.stride_
        shape = (N, Cc // (s * s), H * s, W * s) if self.d2w_ else (N, Cc * s * s, H // s, W // s)
        with torch.cuda.device(x.device):
            out = self._out("top", shape, x)
            call("pcx_dtow", _p(x), _p(out), N, Cc, H, W, s, 1 if self.d2w_ else 0, _stream())
        return [out]


class PseudoQuantOp(_Op):
    """pseudo_quant_opt (extension/pseudo_quant.hpp):
    PseudoQuantOp(channel, bin_num, npart, weight_decay, check_iters, ntop, top_alpha, addr, device, timeit)."""

    def __init__(self, channel, bin_num, npart, weight_decay, check_iters, ntop, top_alpha, addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.channel_, self.bin_num_, self.npart_ = int(channel), int(bin_num), int(npart)
        self.ntop_ = int(ntop)
        self.ctx_ = _resolve(addr)

    def forward(self, x, weight, count, train=False):
        """quant_forward_cuda (pseudo_quant_cuda.cu:157-194).  `train` only drives the training-time weight
        re-centring in the reference (:160, update_weight) and is ignored here (out of scope)."""
        _check_tensor(x)
        _check_tensor(weight, "weight")
        NN, Cc, h, W = x.shape
        with torch.cuda.device(x.device):
            wl = self.ctx_.widths(h, W)
            val = self._out("top0", x.shape, x)
            sym = self._out("top1", x.shape, x) if self.ntop_ > 1 else None
            steps = self._out("steps", (self.channel_, self.bin_num_), x)
            call("pcx_quant_fwd", _p(x), _p(weight), _p(steps), _p(val), _p(sym), None, NN // self.npart_, Cc, h, W,
                 self.npart_, self.bin_num_, int_array(wl), _stream())
        return [val, sym] if sym is not None else [val]


class PseudoDQuantOp(_Op):
    """pseudo_dquant_opt (extension/pseudo_dquant.hpp): PseudoDQuantOp(npart, channel, bin_num, addr, device, timeit)."""

    def __init__(self, npart, channel, bin_num, addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.npart_, self.nchannel_, self.bin_num_ = int(npart), int(channel), int(bin_num)
        self.ctx_ = _resolve(addr)

    def forward(self, x, weight):
        _check_tensor(x)
        _check_tensor(weight, "weight")
        NN, Cc, h, W = x.shape
        if Cc > weight.shape[0]:
            raise PcxError("input has %d channels but the centre table only %d" % (Cc, weight.shape[0]))
        with torch.cuda.device(x.device):
            wl = self.ctx_.widths(h, W)
            out = self._out("top", x.shape, x)
            cen = self._out("centres", (Cc, self.bin_num_), x)
            call("pcx_dquant_fwd", _p(x), _p(weight), _p(cen), _p(out), NN // self.npart_, Cc, h, W, self.npart_,
                 self.bin_num_, int_array(wl), _stream())
        return [out]


# ------------------------------------------------------------------------------------------------ wavefront ops
class _StepOp(_Op):
    """Ops that carry the hidden step counter pidx_ (restart() zeroes it; a shape change zeroes it too)."""

    def __init__(self, device, timeit):
        super().__init__(device, timeit)
        self.pidx_ = 0
        self._shape = None

    def restart(self):
        self.pidx_ = 0

    def _reshape(self, shape):
        shape = tuple(int(s) for s in shape)
        if shape != self._shape:
            self._shape = shape
            self.pidx_ = 0
            return True
        return False

    def _tick(self):
        p = self.pidx_
        self.pidx_ += 1
        return p


class EntropyCtxPadRun2Op(_StepOp):
    """entropy_ctx_pad_run2_opt: EntropyCtxPadRun2Op(pad, npart, ngroup, input, ctx_addr, device, timeit)."""

    def __init__(self, pad, npart, ngroup, input, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.pad_, self.npart_, self.ngroup_, self.input_ = int(pad), int(npart), int(ngroup), bool(input)
        self.ctx_ = _resolve(ctx_addr)

    def forward(self, x):
        _check_tensor(x)
        NN, Cc, hh, ww = x.shape
        h, W = hh - 2 * self.pad_, ww - 2 * self.pad_
        self._reshape((NN, Cc, h, W))
        psum = self._tick()
        if self.input_:
            psum -= 1
        with torch.cuda.device(x.device):
            wl = self.ctx_.widths(h, W)
            band, row, col, tw = self.ctx_.halo(Cc, h, W, self.pad_)
            items, pstart = self.ctx_.pad_items(h, W, self.pad_)
            call("pcx_ctx_pad_step", _p(x), NN // self.npart_, self.npart_, self.ngroup_, Cc // self.ngroup_, h, W, self.pad_,
                 psum, int_array(wl), _p(band), _p(row), _p(col), _p(tw), _p(items),
                 pstart.ctypes.data_as(C.POINTER(C.c_int)), _stream())
        return [x]


class EntropyConv2Op(_StepOp):
    """entropy_conv_opt2: EntropyConv2Op(npart, channel, ngroup, nout, ksize, constrain, pad_in, pad_out, ctx_addr, device, timeit)."""

    def __init__(self, npart, channel, ngroup, nout, kernel_size, constrain, pad_in, pad_out, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.npart_, self.channel_, self.ngroup_, self.nout_ = int(npart), int(channel), int(ngroup), int(nout)
        self.kernel_size_, self.constrain_ = int(kernel_size), int(constrain)
        self.pad_in_, self.pad_out_ = int(pad_in), int(pad_out)
        if self.kernel_size_ != 5:
            raise PcxError("the context model uses 5x5 kernels (got %d)" % self.kernel_size_)
        self.ctx_ = _resolve(ctx_addr)

    def _run(self, x, weight, bias, act, nb):
        _check_tensor(x)
        _check_tensor(weight, "weight")
        _check_tensor(bias, "bias")
        if act is not None:
            _check_tensor(act, "act")
        NN, Cc, hh, ww = x.shape
        h, W = hh - 2 * self.pad_in_, ww - 2 * self.pad_in_
        first = self._reshape((NN, Cc, h, W))
        psum = self._tick()
        num_out = NN // self.npart_
        with torch.cuda.device(x.device):
            out = self._out("top", (NN, self.nout_, h + 2 * self.pad_out_, W + 2 * self.pad_out_), x, zero=True)
            if psum == 0 and not first:
                out.zero_()                                   # cudaMemset at psum == 0 (entropy_conv_cuda_v2.cu:307-309)
            d_order, start = self.ctx_.order(h, W)
            call("pcx_ctx_conv_step", _p(x), _p(weight), _p(bias), _p(act), _p(out), nb, num_out // nb, self.npart_,
                 self.ngroup_, self.channel_ // self.ngroup_, self.nout_ // self.ngroup_, h, W, self.pad_in_, self.pad_out_,
                 self.constrain_, psum, _p(d_order), start.ctypes.data_as(C.POINTER(C.c_int)), _stream())
        return [out]

    def forward(self, x, weight, bias):
        return self._run(x, weight, bias, None, 1)

    def forward_act(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, 1)

    def forward_batch(self, x, weight, bias):
        return self._run(x, weight, bias, None, int(weight.shape[0]))

    def forward_act_batch(self, x, weight, bias, act):
        return self._run(x, weight, bias, act, int(weight.shape[0]))


class EntropyAddOp(_StepOp):
    """entropy_add_opt: EntropyAddOp(npart, channel, ngroup, pad, ctx_addr, device, timeit)."""

    def __init__(self, npart, channel, ngroup, pad, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.npart_, self.channel_, self.ngroup_, self.pad_ = int(npart), int(channel), int(ngroup), int(pad)
        self.ctx_ = _resolve(ctx_addr)

    def forward(self, y, x):
        _check_tensor(y)
        _check_tensor(x, "second input")
        NN, Cc, hh, ww = y.shape
        h, W = hh - 2 * self.pad_, ww - 2 * self.pad_
        self._reshape((NN, Cc, h, W))
        psum = self._tick()
        with torch.cuda.device(y.device):
            d_order, start = self.ctx_.order(h, W)
            call("pcx_ctx_add_step", _p(y), _p(x), NN // self.npart_, self.npart_, self.ngroup_, self.channel_ // self.ngroup_,
                 h, W, self.pad_, psum, _p(d_order), start.ctypes.data_as(C.POINTER(C.c_int)), _stream())
        return [y]


class DInput2Op(_StepOp):
    """d_input_opt2: DInput2Op(nchannel, npart, pad, bias, replicate, ctx_addr, device, timeit)."""

    def __init__(self, nchannel, npart, pad, bias, replicate, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.channel_, self.npart_, self.pad_ = int(nchannel), int(npart), int(pad)
        self.bias_, self.rep_ = float(bias), int(replicate)
        self.ctx_ = _resolve(ctx_addr)

    def forward(self, x):
        _check_tensor(x)
        nimg, _, Hf, W = x.shape
        h = Hf // self.npart_
        self._reshape((nimg * self.npart_, self.channel_, h, W))
        psum = self._tick()
        with torch.cuda.device(x.device):
            out = self._out("top", (self.rep_ * nimg * self.npart_, self.channel_, h + 2 * self.pad_, W + 2 * self.pad_), x)
            d_order, start = self.ctx_.order(h, W)
            call("pcx_dinput_step", _p(x), _p(out), nimg, self.npart_, self.channel_, h, W, self.pad_, C.c_float(self.bias_),
                 self.rep_, psum, _p(d_order), start.ctypes.data_as(C.POINTER(C.c_int)), _stream())
        return [out]


class DExtract2Op(_StepOp):
    """d_extract_opt2: DExtract2Op(npart, nchannel, label, ctx_addr, device, timeit)."""

    def __init__(self, npart, nchannel, label, ctx_addr, device=0, timeit=False):
        super().__init__(device, timeit)
        self.npart_, self.nchannel_, self.label_ = int(npart), int(nchannel), bool(label)
        self.ctx_ = _resolve(ctx_addr)
        self.top_num_ = torch.zeros(1, dtype=torch.int32)

    def _run(self, x, batch, lag):
        _check_tensor(x)
        NN, Cc, h, W = x.shape
        if self._reshape((NN, Cc, h, W)):
            self.top_num_ = torch.zeros(1, dtype=torch.int32)     # d_extract_cuda_v2.cu:18
        psum = self._tick()
        cpn = Cc // self.nchannel_
        nrep = NN // self.npart_
        with torch.cuda.device(x.device):
            out = self._out("top", (nrep, cpn, h * self.npart_, W), x)
            d_order, start = self.ctx_.order(h, W)
            cnt = C.c_int(int(self.top_num_[0]))
            call("pcx_dextract_step", _p(x), _p(out), nrep, self.npart_, self.nchannel_, cpn, h, W, psum, 1 if batch else 0,
                 1 if lag else 0, _p(d_order), start.ctypes.data_as(C.POINTER(C.c_int)), C.byref(cnt), _stream())
            self.top_num_[0] = cnt.value
        return [out, self.top_num_]

    def forward(self, x):
        return self._run(x, False, not self.label_)

    def forward_batch(self, x):
        return self._run(x, True, False)


class EntropyGmmTableOp(_Op):
    """entropy_gmm_table_opt: EntropyGmmTableOp(nstep, bias, num_gaussian, total_region, beta, device, timeit)."""

    def __init__(self, nstep, bias, num_gaussian, total_region, beta=1e-6, device=0, timeit=False):
        super().__init__(device, timeit)
        self.nstep_, self.bias_, self.num_gaussian_ = int(nstep), float(bias), int(num_gaussian)
        self.total_region_, self.beta_ = float(total_region), float(beta)
        if self.num_gaussian_ > 16:
            raise PcxError("the number of Gaussian Distribution in GMM should be less than 16 (extension/entropy_gmm_table_cuda.cu:13)")

    def forward(self, weight, delta, mean, tnum):
        for t in (weight, delta, mean):
            _check_tensor(t)
        tn = int(tnum.reshape(-1)[0])
        rows = weight.shape[0] * weight.shape[2] * weight.shape[3] if weight.dim() == 4 else weight.shape[0]
        with torch.cuda.device(weight.device):
            out = self._out("top", (rows, self.nstep_ + 1), weight)
            call("pcx_gmm_table", _p(weight), _p(delta), _p(mean), tn, self.num_gaussian_, self.nstep_, C.c_float(self.bias_),
                 C.c_float(self.total_region_), C.c_float(self.beta_), 1, _p(out), None, _stream())
        return [out]

    def forward_batch(self, data, tnum):
        """data = cat([logits, delta, mean]) planes (entropy_gmm_table_cuda.cu:155-185); softmax/delta in place."""
        _check_tensor(data)
        tn = int(tnum.reshape(-1)[0])
        total = data.numel()
        stride = total // 3
        with torch.cuda.device(data.device):
            out = self._out("top", (total // 3 // data.shape[1], self.nstep_ + 1), data)
            flat = data.view(-1)
            call("pcx_gmm_table", _p(flat), C.c_void_p(flat.data_ptr() + 4 * stride), C.c_void_p(flat.data_ptr() + 8 * stride),
                 tn, self.num_gaussian_, self.nstep_, C.c_float(self.bias_), C.c_float(self.total_region_),
                 C.c_float(self.beta_), 0, _p(out), None, _stream())
        return [out]


class EntropyGmmOp(_Op):
    """entropy_gmm_opt: EntropyGmmOp(num_gaussian, ignore_label, device, timeit); forward value only."""

    def __init__(self, num_gaussian=3, ignore_label=-1, device=0, timeit=False):
        super().__init__(device, timeit)
        self.num_gaussian_ = int(num_gaussian)

    def forward(self, weight, delta, mean, label):
        for t in (weight, delta, mean, label):
            _check_tensor(t)
        n, ng = weight.shape
        if ng != self.num_gaussian_:
            raise PcxError("the last dim of the weight should be the same as the number of gaussian distributions "
                           "(extension/entropy_gmm_cuda.cu:15)")
        with torch.cuda.device(weight.device):
            out = self._out("top", (n,), weight)
            call("pcx_gmm_nll", _p(weight), _p(delta), _p(mean), _p(label), _p(out), n, ng, _stream())
        return [out]


# ------------------------------------------------------------------------------------------------ out of scope
class _OutOfScope:
    what = ""

    def __init__(self, *a, **k):
        raise NotImplementedError("%s is outside the codec hot path (%s); see DESIGN.md 'Out of scope'" % (type(self).__name__, self.what))


class ProjectsOp(_Op):
    """projects_opt (extension/projects.hpp): ProjectsOp(h_out, w_out, theta[14], phi[14], fov, near, device, timeit) - the 14
    rectilinear viewports on which `pseudo_codec.py --test` measures PSNR / SSIM.  forward only (metric)."""

    def __init__(self, h_out, w_out, theta, phi, fov=0.33333, near=False, device=0, timeit=False):
        super().__init__(device, timeit)
        self.h_out_, self.w_out_, self.fov_, self.near_ = int(h_out), int(w_out), float(fov), bool(near)
        if len(theta) < 14 or len(phi) < 14:
            raise PcxError("ProjectsOp needs 14 viewport angles (extension/projects.hpp:10-13)")
        self.theta_, self.phi_ = [float(v) for v in theta[:14]], [float(v) for v in phi[:14]]
        self._tf = {}

    def _on_device_change(self):
        self._tf.clear()

    def forward(self, x):
        _check_tensor(x)
        N, Cc, H, W = x.shape
        with torch.cuda.device(x.device):
            key = (H, W)
            if key not in self._tf:          # projects_opt::reshape -> update (projects_cuda.cu:153-161)
                tf = torch.zeros((14, self.h_out_ * self.w_out_, 2), dtype=torch.float32, device=x.device)
                call("pcx_project_table", (C.c_float * 14)(*self.theta_), (C.c_float * 14)(*self.phi_), C.c_float(self.fov_),
                     self.h_out_, self.w_out_, H, W, _p(tf), _stream())
                self._tf[key] = tf
            out = self._out("top", (N * 14, Cc, self.h_out_, self.w_out_), x)
            call("pcx_project_fwd", _p(x), _p(self._tf[key]), _p(out), N, Cc, H, W, self.h_out_, self.w_out_, 1 if self.near_ else 0,
                 _stream())
        return [out]

    def backward(self, g):
        raise PcxError("ProjectsOp.backward is training-only and not built (extension/projects_cuda.cu:247-329)")


class ContextReshapeOp(_OutOfScope):
    what = "training loss plumbing, extension/context_reshape_cuda.cu"


class MaskConstrainOp(_OutOfScope):
    what = "training-time weight masking, extension/mask_constrain_cuda.cu"
