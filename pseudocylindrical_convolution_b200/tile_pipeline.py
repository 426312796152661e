"""The pseudocylindrical tile pipeline as ONE fused device path (BASELINE.json config 2):

    SphereSlice -> PseudoPadV2(1) -> nn.Conv2d(Ci, Co, 3) -> [PReLU] -> PseudoFillV2 -> SphereUslice

which in the reference is five operator calls, nine kernels and four full NCHW round trips through memory
(PCONV_operator/SphereSlice.py, PseudoContextV2.py:74-80, model_zoo_v2.py:116-126, SphereUslice.py).  Here it is
three launches per image:

    pcx_slice_pad_nhwc   ERP (C,H,W)  -> halo-padded channels-last band tiles     (HBM-bound gather)
    pcx_conv2d_fwd       tcgen05/TMEM implicit GEMM, TF32 operands, fp32 accumulate; bias, PReLU and the
                         invalid-column fill in the epilogue                        (tensor-pipe bound)
    pcx_uslice_nhwc      band tiles -> ERP (Co,H,W)                                 (HBM-bound gather)

`forward` takes and returns NCHW device tensors exactly like the chain of reference modules would;
`forward_host` is the same call on pinned HOST buffers, streaming image by image with the copies of image
i+1 / i-1 overlapped with the kernels of image i on separate CUDA streams.
"""
import ctypes as C

import torch
from torch import nn

from . import PCONV
from ._lib import ConvDesc, call, int_array
from .PCONV_operator.base import set_weight


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class TilePipeline(nn.Module):
    """slice -> pad(1) -> conv3x3(Ci -> Co) [-> PReLU] -> fill -> uslice on ERP batches (N, Ci, H, W)."""

    def __init__(self, channels_in, channels_out, npart=16, opt=True, act=False, device=0):
        super().__init__()
        self.npart, self.pad = int(npart), 1
        self.gid = int(device)
        self.conv = nn.Conv2d(channels_in, channels_out, 3, 1)          # parameter container; forward() is never called
        self.relu = nn.PReLU(channels_out) if act else None
        weight = set_weight(npart, opt)
        self._slice = PCONV.SphereSliceOp(npart, 0, 0, weight, device, False)
        self._uslice = PCONV.SphereUsliceOp(npart, 0, 0, weight, device, False)
        self._ctx = PCONV.PseudoContextOp(npart, 20, weight, device, False)
        self._bufs = {}
        self._streams = None
        self.to(torch.device("cuda", self.gid))

    # ---------------------------------------------------------------------------------------------- geometry
    def _geometry(self, H, W, like):
        wl, s_src, s_wt = self._slice._geometry(H, W, like, "pcx_slice_table")
        _, u_src, u_wt = self._uslice._geometry(H, W, like, "pcx_uslice_table")
        h = H // self.npart
        band, row, col, tw = self._ctx.halo(self.conv.in_channels, h, W, self.pad)
        return wl, (s_src, s_wt), (u_src, u_wt), (band, row, col, tw)

    def _tiles(self, key, shape, dev):
        buf = self._bufs.get(key)
        if buf is None or tuple(buf.shape) != tuple(shape) or buf.device != dev:
            buf = torch.zeros(shape, dtype=torch.float32, device=dev)
            self._bufs[key] = buf
        return buf

    # ---------------------------------------------------------------------------------------------- one image
    def _run_image(self, x, out, slot=0):
        """x (n, Ci, H, W), out (n, Co, H, W): device tensors; all launches go to the current stream."""
        n, Ci, H, W = x.shape
        Co = self.conv.out_channels
        if H % self.npart != 0:
            raise PCONV.PcxError("height should be multipler of the number of parts (extension/math_cuda.cu:179)")
        h, p = H // self.npart, self.pad
        wl, (s_src, s_wt), (u_src, u_wt), (band, row, col, tw) = self._geometry(H, W, x)
        planes = n * self.npart
        xt = self._tiles(("xt", slot), (planes, h + 2 * p, W + 2 * p, Ci), x.device)
        yt = self._tiles(("yt", slot), (planes, h, W, Co), x.device)
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        # the tile buffers are zero-initialised once and the band geometry is fixed, so the columns beyond the
        # bands never need rewriting (zero_invalid = 0)
        call("pcx_slice_pad_nhwc", _p(x), _p(xt), n, Ci, H, W, self.npart, p, int_array(wl), _p(s_src), _p(s_wt),
             _p(band), _p(row), _p(col), _p(tw), W + 2 * p, 0, s)
        d = ConvDesc()
        d.N, d.npart = n, self.npart
        d.Ci, d.Hi, d.in_pitch = Ci, h + 2 * p, W + 2 * p
        d.Co, d.Ho, d.Wo = Co, h, W
        d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = h, W, 0, 0
        d.k, d.stride = 3, 1
        d.act = 1 if self.relu is not None else 0
        d.impl = 2
        d.aux_rows, d.aux_pitch, d.aux_y0, d.aux_x0 = h, W, 0, 0
        for g in range(self.npart):
            d.wl_out[g] = wl[g]
        call("pcx_conv2d_fwd", C.byref(d), _p(xt), _p(self._packed_weights(s)), _p(self.conv.bias.data),
             _p(self.relu.weight.data) if self.relu is not None else None, None, None, _p(yt), s)
        call("pcx_uslice_nhwc", _p(yt), _p(out), n, Co, h, W, self.npart, h, W, 0, 0, int_array(wl), _p(u_src), _p(u_wt), s)

    def _packed_weights(self, stream):
        """Weights in the tap-major TF32 layout of the tensor-core kernel; repacked only when the parameter changes."""
        w = self.conv.weight
        key = (w.data_ptr(), w._version, tuple(w.shape))
        if self._bufs.get("wkey") != key:
            Co, Ci, k, _ = w.shape
            n = call("pcx_conv_pack_weights", None, None, Co, Ci, k, stream)
            packed = torch.empty(n, dtype=torch.float32, device=w.device)
            call("pcx_conv_pack_weights", _p(w.data), _p(packed), Co, Ci, k, stream)
            self._bufs["wkey"], self._bufs["wpacked"] = key, packed
        return self._bufs["wpacked"]

    LAUNCHES_PER_IMAGE = 3      # slice_pad, conv, uslice

    # ---------------------------------------------------------------------------------------------- public
    @torch.no_grad()
    def forward(self, x, out=None, images_per_launch=1):
        """x: (N, Ci, H, W) float32 CUDA tensor -> (N, Co, H, W)."""
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise TypeError("input must be a contiguous float32 CUDA tensor")
        N, _, H, W = x.shape
        if out is None:
            out = torch.empty((N, self.conv.out_channels, H, W), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            for i in range(0, N, images_per_launch):
                self._run_image(x[i:i + images_per_launch], out[i:i + images_per_launch])
        return out

    @torch.no_grad()
    def forward_host(self, x_host, out_host):
        """x_host / out_host: sequences of pinned CPU tensors (Ci,H,W) / (Co,H,W), one per image.  Copies in, runs and
        copies out image by image on three streams with two device slots; returns after everything has landed."""
        dev = torch.device("cuda", self.gid)
        with torch.cuda.device(dev):
            if self._streams is None:
                self._streams = tuple(torch.cuda.Stream(dev) for _ in range(3))
            s_in, s_run, s_out = self._streams
            Ci, H, W = x_host[0].shape
            Co = self.conv.out_channels
            d_in = [self._tiles(("din", k), (1, Ci, H, W), dev) for k in range(2)]
            d_out = [self._tiles(("dout", k), (1, Co, H, W), dev) for k in range(2)]
            in_free = [None, None]
            out_free = [None, None]
            start = torch.cuda.Event()
            start.record(torch.cuda.current_stream())
            for st in self._streams:
                st.wait_event(start)
            for i, (xh, oh) in enumerate(zip(x_host, out_host)):
                k = i & 1
                with torch.cuda.stream(s_in):
                    if in_free[k] is not None:
                        s_in.wait_event(in_free[k])
                    d_in[k][0].copy_(xh, non_blocking=True)
                    ready = torch.cuda.Event()
                    ready.record(s_in)
                with torch.cuda.stream(s_run):
                    s_run.wait_event(ready)
                    if out_free[k] is not None:
                        s_run.wait_event(out_free[k])
                    self._run_image(d_in[k], d_out[k], slot=0)
                    in_free[k] = torch.cuda.Event()
                    in_free[k].record(s_run)
                    done = torch.cuda.Event()
                    done.record(s_run)
                with torch.cuda.stream(s_out):
                    s_out.wait_event(done)
                    oh.copy_(d_out[k][0], non_blocking=True)
                    out_free[k] = torch.cuda.Event()
                    out_free[k].record(s_out)
            cur = torch.cuda.current_stream()
            for st in self._streams:
                e = torch.cuda.Event()
                e.record(st)
                cur.wait_event(e)
        return out_host


def reference_chain(x, conv, relu, npart=16, opt=True, device=0, impl=1):
    """The same computation through the separate operator modules (the reference's own module sequence), used by the
    tests as the on-device comparison: NCHW throughout, fp32 CUDA-core convolution (impl=1)."""
    from .PCONV_operator import PseudoContextV2, PseudoPadV2, SphereSlice, SphereUslice
    from .model_zoo_v2 import pconv
    sl = SphereSlice(npart, pad=0, opt=opt, device=device)
    us = SphereUslice(npart, pad=0, opt=opt, device=device)
    ctx = PseudoContextV2(npart, opt, device=device)
    pad = PseudoPadV2(1, npart, ctx, device=device)
    t = sl(x)
    h, W = t.shape[2:]
    wl = ctx.op[device].widths(h, W)
    y = pconv(pad(t), conv, npart, wl, prelu=relu, impl=impl)
    return us(y)
