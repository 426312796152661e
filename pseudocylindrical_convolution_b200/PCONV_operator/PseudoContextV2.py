"""Tile-seam operators of the transforms (reference: PCONV_operator/PseudoContextV2.py).

PseudoContextV2 / PseudoEntropyContext own the band geometry and halo tables; PseudoPadV2, PseudoFillV2,
PseudoEntropyPad, PseudoQUANTV2, PseudoDQUANT and PseudoGDNV2 hold a native op bound to such a context.
Parameter names match the reference so its checkpoints load unchanged.
"""
import math

import torch
from torch import nn

from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule
from .EntropyContextNew import EntropyContextNew
from .base import set_weight


class _ContextModule(BaseOpModule):
    def setup_context(self, w):
        for op in self.op.values():
            op.start_context(w)

    def get_addr(self, gid):
        return self.op[gid].addr()


class PseudoContextV2(_ContextModule):

    def __init__(self, npart, opt=True, rt=20, device=0, time_it=False):
        super().__init__(device)
        weight = set_weight(npart, opt)
        self.op = {gid: PCONV.PseudoContextOp(npart, rt, weight, gid, time_it) for gid in self.device_list}

    def produce_fill_param(self, gid, h, w):
        return self.op[gid].produce_fill_param(h, w)


class PseudoEntropyContext(_ContextModule):

    def __init__(self, npart, context_version=1, opt=True, rt=20, device=0, time_it=False):
        super().__init__(device)
        weight = set_weight(npart, opt)
        self.op = {gid: PCONV.PseudoEntropyContextOp(npart, rt, context_version, weight, gid, time_it) for gid in self.device_list}


class PseudoPadV2(BaseOpModule):
    """(N*npart, C, h, W) -> (N*npart, C, h+2p, W+2p): interior copy, interpolated halo rows from the
    neighbouring bands (mirrored + shifted by 180 degrees across the poles), longitude wrap."""

    def __init__(self, pad, npart, ctx: PseudoContextV2, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.PseudoPadOp(pad, npart, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]


class PseudoEntropyPad(BaseOpModule):
    """Causal variant used by the (training-time, full-tensor) context model: future samples are zero."""

    def __init__(self, pad, npart, ctx: PseudoEntropyContext, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.PseudoEntropyPadOp(pad, npart, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]


class PseudoFillV2(BaseOpModule):
    """In-place: write `fvalue` outside the valid columns of every band (and inside the pad-trim border)."""

    def __init__(self, pad, npart, ctx, fvalue=0, trim=0, device=0, time_it=False):
        super().__init__(device)
        version = 0 if isinstance(ctx, PseudoContextV2) else (1 if isinstance(ctx, PseudoEntropyContext) else 2)
        self.op = {gid: PCONV.PseudoFillOp(pad, npart, fvalue, trim, ctx.get_addr(gid), version, gid, time_it)
                   for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]


class PseudoGDNV2(nn.Module):
    """Generalised divisive normalisation restricted to the valid columns:
    y_i = x_i / sqrt(beta_i + sum_j gamma_ij x_j^2)   (inverse=True multiplies instead)."""

    def __init__(self, ch, npart, ctx: PseudoContextV2, device=0, inverse=False, beta_min=1e-6, gamma_init=.1,
                 reparam_offset=2 ** -18):
        super().__init__()
        self.inverse = inverse
        self.beta_min = beta_min
        self.gamma_init = gamma_init
        self.reparam_offset = float(reparam_offset)
        self.npart = npart
        self.ctx = [ctx]                       # not a submodule: the context is owned by the codec
        gid = device if isinstance(device, int) else device[0]
        dev = torch.device("cuda:%d" % gid)
        pedestal = torch.FloatTensor([reparam_offset]) ** 2
        self.beta = nn.Parameter(torch.sqrt(torch.ones(ch) + pedestal).to(dev))
        self.gamma = nn.Parameter(torch.sqrt(self.gamma_init * torch.eye(ch) + pedestal).to(dev))
        self._eff = None

    def _effective(self):
        """beta', gamma' (LowerBound + square - pedestal) computed on the device, cached until the parameters change."""
        import ctypes as C
        from .._lib import call
        key = (self.beta._version, self.gamma._version, self.beta.data_ptr(), self.gamma.data_ptr())
        if self._eff is None or self._eff[0] != key:
            be = torch.empty_like(self.beta)
            ge = torch.empty_like(self.gamma)
            call("pcx_gdn_params", C.c_void_p(self.beta.data_ptr()), C.c_void_p(self.gamma.data_ptr()),
                 C.c_void_p(be.data_ptr()), C.c_void_p(ge.data_ptr()), self.beta.numel(), self.beta_min,
                 self.reparam_offset, C.c_void_p(torch.cuda.current_stream().cuda_stream))
            self._eff = (key, be, ge)
        return self._eff[1], self._eff[2]

    def forward(self, inputs, residual=None):
        """residual (optional, same shape): returns fill(residual + gdn(inputs)) in the same pass."""
        import ctypes as C
        from .._lib import call, int_array
        x = contiguous(inputs)
        NN, ch, h, W = x.shape
        with torch.cuda.device(x.device):
            be, ge = self._effective()
            wl = self.ctx[0].op[x.device.index].widths(h, W)
            y = torch.empty_like(x)
            call("pcx_gdn_fwd", C.c_void_p(x.data_ptr()), C.c_void_p(be.data_ptr()), C.c_void_p(ge.data_ptr()),
                 C.c_void_p(contiguous(residual).data_ptr()) if residual is not None else None,
                 C.c_void_p(y.data_ptr()), NN // self.npart, ch, h, W, self.npart, int_array(wl), 1 if self.inverse else 0,
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return y


class PseudoQUANTV2(BaseOpModule):
    """Learned 8-level scalar quantiser per channel; returns (dequantised, symbols) when ntop == 2."""

    def __init__(self, channel, bin_num, npart, ctx: PseudoContextV2, check_iters=100, weight_decay=0.9, ntop=1,
                 top_alpha=0.1, device_id=0, time_flag=False):
        super().__init__(device_id)
        dev = "cuda:%d" % self.device_list[0]
        first = 1. / (bin_num + 1)
        w = torch.full((channel, bin_num), math.log(first), dtype=torch.float32)
        w[:, 0] = first
        self.weight = nn.Parameter(w.to(dev))
        self.count = nn.Parameter(torch.zeros((channel, bin_num), dtype=torch.float32).to(dev))
        self.op = {gid: PCONV.PseudoQuantOp(channel, bin_num, npart, weight_decay, check_iters, ntop, top_alpha,
                                            ctx.get_addr(gid), gid, time_flag) for gid in self.device_list}

    def forward(self, x):
        out = self.native(x).forward(contiguous(x), self.weight.data, self.count.data, self.training)
        return out[0] if len(out) == 1 else (out[0], out[1])


class PseudoDQUANT(BaseOpModule):
    """Symbols -> reconstruction centres (running sum of the exponentiated steps)."""

    def __init__(self, channel, bin_num, npart, ctx: PseudoContextV2, device_id=0, time_flag=False):
        super().__init__(device_id)
        dev = "cuda:%d" % self.device_list[0]
        self.weight = nn.Parameter(torch.zeros((channel, bin_num), dtype=torch.float32).to(dev))
        self.op = {gid: PCONV.PseudoDQuantOp(npart, channel, bin_num, ctx.get_addr(gid), gid, time_flag) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x), self.weight.data)[0]
