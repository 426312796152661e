"""Channels-last execution of the analysis / synthesis transforms on the tcgen05 convolution kernels.

The reference runs every layer as PseudoPadV2 -> nn.Conv2d (cuDNN) -> PReLU / Sigmoid / add / GDN -> PseudoFillV2 on
NCHW tensors (model_zoo_v2.py:36-211): ~5 full passes over the activation per layer.  Here an activation is a band-tile
buffer [plane][h + 4][W + 4][C] (channels last, halo of 2 on every side, zero beyond the band width) and a layer is

    pcx_conv2d_fwd (TMA -> tcgen05.mma -> TMEM; bias, PReLU / sigmoid / 1/sqrt, gate, residual and the invalid-column fill in
                    the epilogue, written straight into the INTERIOR of the next padded buffer)
    pcx_halo_fill_nhwc (in-place halo refresh = PseudoPadV2 of that buffer)

A consumer that needs pad p <= 2 reads the sub-window starting at (2 - p, 2 - p): the inner ring of a pad-2 halo equals the
pad-1 halo (same source rows, same wrap).  GDN / IGDN are a 1x1 tensor-core convolution of x^2 by gamma' with bias beta' and
an `x * rsqrt(.)` / `x * sqrt(.)` epilogue (PseudoContextV2.py:186-216); depth-to-space (Dtow) writes into the next padded
buffer.  Parameters are read from the same nn.Modules as the NCHW path, so checkpoints are shared.
"""
import ctypes as C

import torch

from ._lib import ConvDesc, call, int_array

HALO = 2
ACT_NONE, ACT_PRELU, ACT_SIGMOID, ACT_RSQRT, ACT_SQRT = 0, 1, 2, 3, 4


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class View:
    """A window (y0, x0, h, W) into a channels-last plane buffer [planes][rows][pitch][C]."""
    __slots__ = ("buf", "rows", "pitch", "y0", "x0", "h", "W", "C", "_pool", "_key")

    def __init__(self, buf, y0, x0, h, W):
        self.buf = buf
        _, self.rows, self.pitch, self.C = buf.shape
        self.y0, self.x0, self.h, self.W = y0, x0, h, W

    @property
    def planes(self):
        return self.buf.shape[0]

    def ptr(self):
        return C.c_void_p(self.buf.data_ptr() + 4 * ((self.y0 * self.pitch + self.x0) * self.C))

    def padded(self, p):
        """the same tile with a halo of p (requires a padded buffer whose halo has been refreshed)"""
        assert self.y0 >= p and self.x0 >= p
        return View(self.buf, self.y0 - p, self.x0 - p, self.h + 2 * p, self.W + 2 * p)


class Runner:
    """Owns the buffers, packed weights and geometry of one transform instance on one device."""

    def __init__(self, ctx_op, npart):
        self.ctx = ctx_op              # PCONV.PseudoContextOp: widths + halo tables
        self.npart = npart
        self._packed = {}
        self._pool = {}
        self._live = []

    # ------------------------------------------------------------------------------------------ buffers
    def _buffer(self, shape, zero):
        """Buffers are pooled per exact shape.  Padded buffers are zero-initialised ONCE: later passes rewrite the
        interior and the halo ring only, the cells beyond the bands stay zero."""
        key = (tuple(shape), zero)
        free = self._pool.setdefault(key, [])
        for i, b in enumerate(free):
            if not any(b is v for v in self._live):
                return b
        b = (torch.zeros if zero else torch.empty)(shape, dtype=torch.float32, device=self.device)
        free.append(b)
        return b

    def padded_act(self, planes, h, W, Cc):
        b = self._buffer((planes, h + 2 * HALO, W + 2 * HALO, Cc), True)
        self._live.append(b)
        return View(b, HALO, HALO, h, W)

    def plain(self, planes, h, W, Cc):
        b = self._buffer((planes, h, W, Cc), False)
        self._live.append(b)
        return View(b, 0, 0, h, W)

    def release(self, *views):
        for v in views:
            self._live = [b for b in self._live if b is not v.buf]

    def reset(self, device, problem=None):
        """problem = (batch, height, width) of the pass that starts; a different problem drops the pooled buffers of the
        previous one (the pool is keyed by exact shape and would otherwise grow with every new image size)."""
        self.device = device
        self._live = []
        if problem is not None and problem != getattr(self, "_problem", None):
            self._pool.clear()
            self._problem = problem

    # ------------------------------------------------------------------------------------------ parameters
    def packed(self, weight, ci_pad=None, d2w=False):
        """OIHW parameter -> tap-major TF32 layout of the tensor-core kernels, cached until the parameter changes.
        d2w: output channels in the depth-to-space order of pcx_conv2d_fwd impl 3."""
        key = (id(weight), d2w)
        ver = (weight.data_ptr(), weight._version, tuple(weight.shape), ci_pad)
        hit = self._packed.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        w = weight.detach()
        if w.dim() == 2:
            w = w.view(w.shape[0], w.shape[1], 1, 1)
        Co, Ci, k, _ = w.shape
        if ci_pad is not None and ci_pad > Ci:
            wp = torch.zeros((Co, ci_pad, k, k), dtype=w.dtype, device=w.device)
            wp[:, :Ci] = w
            w, Ci = wp, ci_pad
        w = w.contiguous()
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        fn = "pcx_conv_pack_weights_d2w" if d2w else "pcx_conv_pack_weights"
        n = call(fn, None, None, Co, Ci, k, s)
        out = torch.empty(n, dtype=torch.float32, device=w.device)
        call(fn, _p(w), _p(out), Co, Ci, k, s)
        self._packed[key] = (ver, out)
        return out

    def widths(self, h, W):
        return self.ctx.widths(h, W)

    # ------------------------------------------------------------------------------------------ ops
    def conv(self, x: View, weight, bias, k, stride, Co, out: View, wl_out, act=ACT_NONE, slope=None, mul: View = None,
             residual: View = None, ci_pad=None, d2w=False, square_input=False):
        """d2w: Dtow(2) fused into the store - `out` is the (2 Ho, 2 Wo, Co / 4) view the pixel shuffle would produce.
        square_input: the convolution reads x * x, squared on chip (GDN)."""
        Ho, Wo = (x.h - k) // stride + 1, (x.W - k) // stride + 1
        if d2w:
            assert (2 * Ho, 2 * Wo) == (out.h, out.W) and 4 * out.C == Co and mul is None and residual is None
        else:
            assert (Ho, Wo) == (out.h, out.W) and out.C == Co, ((Ho, Wo), (out.h, out.W), out.C, Co)
        d = ConvDesc()
        d.N, d.npart = x.planes // self.npart, self.npart
        d.Ci, d.Hi, d.in_pitch, d.in_plane_rows = x.C, x.rows - x.y0, x.pitch, x.rows
        d.Co, d.Ho, d.Wo = Co, Ho, Wo
        d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = out.rows, out.pitch, out.y0, out.x0
        d.k, d.stride, d.act, d.impl = k, stride, act, 3 if d2w else 2
        d.square_input = 1 if square_input else 0
        aux = mul if mul is not None else residual
        if aux is not None:
            if mul is not None and residual is not None:
                assert (mul.rows, mul.pitch, mul.y0, mul.x0) == (residual.rows, residual.pitch, residual.y0, residual.x0)
            assert aux.C == Co and (aux.h, aux.W) == (Ho, Wo)
            d.aux_rows, d.aux_pitch, d.aux_y0, d.aux_x0 = aux.rows, aux.pitch, aux.y0, aux.x0
        else:
            d.aux_rows, d.aux_pitch = Ho, Wo
        for g in range(self.npart):
            d.wl_out[g] = min(int(wl_out[g]), Wo)
        call("pcx_conv2d_fwd", C.byref(d), x.ptr(), _p(self.packed(weight, ci_pad, d2w)), _p(bias), _p(slope),
             _p(mul.buf) if mul is not None else None, _p(residual.buf) if residual is not None else None, _p(out.buf),
             C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out

    def conv_m(self, x, conv, out, wl_out, prelu=None, sigmoid=False, mul=None, residual=None, ci_pad=None, d2w=False):
        """convolution described by an nn.Conv2d parameter container (+ optional nn.PReLU)"""
        act = ACT_PRELU if prelu is not None else (ACT_SIGMOID if sigmoid else ACT_NONE)
        return self.conv(x, conv.weight, conv.bias.data if conv.bias is not None else None, conv.kernel_size[0], conv.stride[0],
                         conv.out_channels, out, wl_out, act, prelu.weight.data if prelu is not None else None, mul, residual, ci_pad, d2w)

    def halo(self, a: View):
        """PseudoPadV2(2) of a padded activation, in place"""
        assert a.y0 == HALO and a.x0 == HALO
        wl = self.widths(a.h, a.W)
        band, row, col, tw = self.ctx.halo(a.C, a.h, a.W, HALO)
        call("pcx_halo_fill_nhwc", _p(a.buf), a.planes // self.npart, a.C, a.h, a.W, self.npart, HALO, int_array(wl), _p(band), _p(row),
             _p(col), _p(tw), a.rows, a.pitch, a.y0, a.x0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return a

    def dtow(self, x: View, out: View):
        assert x.C == 4 * out.C and (out.h, out.W) == (2 * x.h, 2 * x.W)
        call("pcx_dtow_nhwc", _p(x.buf), _p(out.buf), x.planes, out.C, x.h, x.W, x.rows, x.pitch, x.y0, x.x0, out.rows, out.pitch,
             out.y0, out.x0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return out

    def gdn(self, gdn_module, z: View, residual: View, out: View, wl):
        """out = fill(residual + z / sqrt(beta' + gamma' z^2))   (inverse: z * sqrt(.)) - PseudoContextV2.py:186-216"""
        assert z.y0 == 0 and z.x0 == 0 and z.rows == z.h and z.pitch == z.W
        be, ge = gdn_module._effective()
        # the 1x1 GEMM over z^2: z is squared inside the kernel's shared-memory pipeline (no x^2 tensor, no square kernel);
        # the same z is the epilogue's gate operand
        self.conv(z, ge, be, 1, 1, z.C, out, wl, ACT_SQRT if gdn_module.inverse else ACT_RSQRT, None, z, residual, square_input=True)
        return out


# ---------------------------------------------------------------------------------------------------- blocks
def _residual_block(R: Runner, m, X: View):
    """ResidualBlock (model_zoo_v2.py:36-53): x + conv1x1(PReLU(conv3x3(PReLU(conv1x1(pad1(x))))))"""
    wl = R.widths(X.h, X.W)
    ch = m.conv1.out_channels
    y1 = R.conv_m(X.padded(1), m.conv1, R.plain(X.planes, X.h + 2, X.W + 2, ch), [w + 2 for w in wl], prelu=m.relu1)
    y2 = R.conv_m(y1, m.conv2, R.plain(X.planes, X.h, X.W, ch), wl, prelu=m.relu2)
    out = R.conv_m(y2, m.conv3, R.padded_act(X.planes, X.h, X.W, X.C), wl, residual=X)
    R.release(y1, y2)
    return R.halo(out)


def _attention_block(R: Runner, m, X: View):
    """AttentionBlock (model_zoo_v2.py:55-76): x + trunk(x) * sigmoid(conv1x1(attention(x)))"""
    wl = R.widths(X.h, X.W)
    t = X
    for i in range(3):
        n = _residual_block(R, m.trunk[i], t)
        if t is not X:
            R.release(t)
        t = n
    a = X
    for i in range(3):
        n = _residual_block(R, m.attention[i], a)
        if a is not X:
            R.release(a)
        a = n
    out = R.conv_m(a, m.attention[3], R.padded_act(X.planes, X.h, X.W, X.C), wl, sigmoid=True, mul=t, residual=X)
    R.release(t, a)
    return R.halo(out)


def _residual_block_v2(R: Runner, m, X: View):
    """ResidualBlockV2 (model_zoo_v2.py:78-93): x + PReLU(conv3x3(PReLU(conv3x3(pad2(x)))))"""
    wl = R.widths(X.h, X.W)
    y = R.conv_m(X.padded(2), m.conv1, R.plain(X.planes, X.h + 2, X.W + 2, X.C), [w + 2 for w in wl], prelu=m.relu1)
    out = R.conv_m(y, m.conv2, R.padded_act(X.planes, X.h, X.W, X.C), wl, prelu=m.relu2, residual=X)
    R.release(y)
    return R.halo(out)


def _residual_block_down(R: Runner, m, X: View, ci_pad=None):
    """ResidualBlockDown (model_zoo_v2.py:95-114): conv1x1/2(x) + GDN(conv3x3(pad1(PReLU(conv3x3/2(pad1(x))))))"""
    h2, W2 = X.h // 2, X.W // 2
    wl2 = R.widths(h2, W2)
    Co = m.conv1.out_channels
    t = R.conv_m(X, m.short_cut, R.plain(X.planes, h2, W2, Co), wl2, ci_pad=ci_pad)
    y = R.halo(R.conv_m(X.padded(1), m.conv1, R.padded_act(X.planes, h2, W2, Co), wl2, prelu=m.relu1, ci_pad=ci_pad))
    z = R.conv_m(y.padded(1), m.conv2, R.plain(X.planes, h2, W2, Co), wl2)
    out = R.gdn(m.relu2, z, t, R.padded_act(X.planes, h2, W2, Co), wl2)
    R.release(t, y, z)
    return R.halo(out)


def _residual_block_up(R: Runner, m, X: View):
    """ResidualBlockUp (model_zoo_v2.py:153-175): d2w(conv1x1(x)) + IGDN(conv3x3(pad1(d2w(PReLU(conv3x3(pad1(x)))))))"""
    wl = R.widths(X.h, X.W)
    h2, W2 = 2 * X.h, 2 * X.W
    wl2 = R.widths(h2, W2)
    assert all(int(b) <= 2 * int(a) for a, b in zip(wl, wl2)), "widths that do not double take the NCHW path (DecoderV2.widths_double)"
    ch = X.C
    fused = 4 * ch == m.conv1.out_channels and ch in (96, 192)      # Dtow folded into the convolutions' stores (impl 3)
    if fused:
        b1u = R.halo(R.conv_m(X.padded(1), m.conv1, R.padded_act(X.planes, h2, W2, ch), wl, prelu=m.relu1, d2w=True))
    else:
        b1 = R.conv_m(X.padded(1), m.conv1, R.plain(X.planes, X.h, X.W, 4 * ch), wl, prelu=m.relu1)
        b1u = R.halo(R.dtow(b1, R.padded_act(X.planes, h2, W2, ch)))
        R.release(b1)
    z = R.conv_m(b1u.padded(1), m.conv2, R.plain(X.planes, h2, W2, ch), wl2)
    if fused:
        su = R.conv_m(X, m.short_cut, R.plain(X.planes, h2, W2, ch), wl, d2w=True)
    else:
        s = R.conv_m(X, m.short_cut, R.plain(X.planes, X.h, X.W, 4 * ch), wl)
        su = R.dtow(s, R.plain(X.planes, h2, W2, ch))
        R.release(s)
    R.release(b1u)
    out = R.gdn(m.relu2, z, su, R.padded_act(X.planes, h2, W2, ch), wl2)
    R.release(z, su)
    return R.halo(out)


def _sphere_conv2(R: Runner, m, X: View):
    """SphereConv2 (model_zoo_v2.py:116-126): fill(conv3x3/2(pad1(x)))"""
    h2, W2 = X.h // 2, X.W // 2
    out = R.conv_m(X.padded(1), m.conv, R.padded_act(X.planes, h2, W2, m.conv.out_channels), R.widths(h2, W2))
    return R.halo(out)


def _run_blocks(R, blocks, X, first_ci_pad=None):
    from . import model_zoo_v2 as mz
    table = {mz.ResidualBlockDown: _residual_block_down, mz.ResidualBlockV2: _residual_block_v2, mz.AttentionBlock: _attention_block,
             mz.SphereConv2: _sphere_conv2, mz.ResidualBlockUp: _residual_block_up}
    for i, blk in enumerate(blocks):
        fn = table[type(blk)]
        n = fn(R, blk, X, first_ci_pad) if (i == 0 and first_ci_pad) else fn(R, blk, X)
        R.release(X)
        X = n
    return X


# ---------------------------------------------------------------------------------------------------- CUDA graphs
# A transform is ~100 launches; at 512x1024 (BASELINE configs[0]) each kernel runs for 5-20 us and the pass is bound by the
# host issuing them through ctypes (2.4 ms for 102 launches, profiles/r1w_codec_bench.jsonl).  The second consecutive call
# with the same (shape, device, parameter versions) captures the whole pass - our launches, the tensor-map encodes baked
# into their parameters, and the few torch elementwise kernels - into one CUDA graph; later calls copy the input into the
# graph's static buffer and replay it.  The graph is dropped as soon as the problem or any parameter changes (the Runner
# recycles its tile buffers then), so a stale capture can never run.
def _param_stamp(module):
    return sum(int(p._version) for p in module.parameters()) + sum(int(b._version) for b in module.buffers())


def _run_graphed(owner, fn, x):
    from . import config
    if not getattr(config, "CUDA_GRAPHS", True):
        return fn(x)
    st = owner.__dict__.setdefault("_pcx_graph", {"key": None, "seen": None, "graph": None, "failed": False})
    key = (tuple(x.shape), x.device.index, x.data_ptr() % 16, _param_stamp(owner))
    if st["graph"] is not None and st["key"] == key:
        st["x"].copy_(x)
        st["graph"].replay()
        return st["y"].clone()
    st["graph"] = st["x"] = st["y"] = st["key"] = None
    y = fn(x)                                   # eager pass (also the warm-up of the capture below)
    if st["seen"] == key and not st["failed"]:
        try:
            sx = x.clone()
            torch.cuda.synchronize(x.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                sy = fn(sx)
            st.update(graph=g, x=sx, y=sy, key=key)
        except Exception as e:                  # capture is an optimisation: keep the eager path if the driver refuses
            st["failed"] = True
            import warnings
            warnings.warn("pcx: CUDA-graph capture of the transform failed (%r); running eagerly" % (e,))
    st["seen"] = key
    return y


def encoder_forward(enc, erp, slice_op):
    """EncoderV2 + SphereSlice on an ERP batch; graph-replayed from the third call with the same problem (see above).
    Runs with the input's GPU as the current device (streams, allocations and launches follow the tensor, not the caller)."""
    with torch.cuda.device(erp.device):
        return _run_graphed(enc, lambda t: _encoder_forward(enc, t, slice_op), erp)


def decoder_forward(dec, code, uslice_op):
    """DecoderV2 + SphereUslice; graph-replayed from the third call with the same problem (see above)."""
    with torch.cuda.device(code.device):
        return _run_graphed(dec, lambda t: _decoder_forward(dec, t, uslice_op), code)


@torch.no_grad()
def _encoder_forward(enc, erp, slice_op):
    """EncoderV2 (model_zoo_v2.py:129-151) on an ERP batch (N, 3, H, W) -> code (N*npart, code_channels, H/16/npart, W/16), NCHW."""
    R = enc._runner()
    R.reset(erp.device, tuple(erp.shape))
    npart = enc.npart
    N, Cin, H, W = erp.shape
    h = H // npart
    # the first layer's 3 input channels are zero-padded to one 32-channel K block
    x32 = torch.zeros((N, 32, H, W), dtype=torch.float32, device=erp.device)
    x32[:, :Cin] = erp
    X = R.padded_act(N * npart, h, W, 32)
    wl, src, wt = slice_op._geometry(H, W, erp, "pcx_slice_table")
    band, row, col, tw = R.ctx.halo(32, h, W, HALO)
    call("pcx_slice_pad_nhwc", _p(x32), _p(X.buf), N, 32, H, W, npart, HALO, int_array(wl), _p(src), _p(wt), _p(band), _p(row), _p(col),
         _p(tw), X.pitch, 0, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    X = _run_blocks(R, [enc.net[i] for i in range(9)], X, first_ci_pad=32)
    code = R.conv_m(X, enc.net[9], R.plain(X.planes, X.h, X.W, enc.net[9].out_channels), R.widths(X.h, X.W), sigmoid=True)
    out = code.buf.permute(0, 3, 1, 2).contiguous()
    R.reset(erp.device)
    return out


@torch.no_grad()
def _decoder_forward(dec, code, uslice_op):
    """DecoderV2 (model_zoo_v2.py:189-211) + SphereUslice: code (N*npart, C, h, w) NCHW -> ERP (N, 3, 16 h npart, 16 w)."""
    R = dec._runner()
    R.reset(code.device, tuple(code.shape))
    npart = dec.npart
    planes, Cc, h, W = code.shape
    x = R.plain(planes, h, W, Cc)
    x.buf.copy_(code.permute(0, 2, 3, 1))
    X = R.halo(R.conv_m(x, dec.net[0].conv, R.padded_act(planes, h, W, dec.net[0].conv.out_channels), R.widths(h, W)))
    R.release(x)
    X = _run_blocks(R, [dec.net[i] for i in range(1, 10)], X)
    last = dec.net[11]
    y = R.conv_m(X.padded(1), last, R.plain(planes, X.h, X.W, last.out_channels), R.widths(X.h, X.W))
    img = R.dtow(y, R.plain(planes, 2 * X.h, 2 * X.W, last.out_channels // 4))
    N = planes // npart
    Hh, Ww = img.h, img.W
    erp = torch.empty((N, img.C, Hh * npart, Ww), dtype=torch.float32, device=code.device)
    wl, src, wt = uslice_op._geometry(Hh * npart, Ww, erp, "pcx_uslice_table")
    call("pcx_uslice_nhwc", _p(img.buf), _p(erp), N, img.C, Hh, Ww, npart, img.rows, img.pitch, 0, 0, int_array(wl), _p(src), _p(wt),
         C.c_void_p(torch.cuda.current_stream().cuda_stream))
    R.reset(code.device)
    return erp
