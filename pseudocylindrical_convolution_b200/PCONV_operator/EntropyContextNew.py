"""Wavefront (codec-time) context-model operators (reference: PCONV_operator/EntropyContextNew.py).

All of them evaluate only the cells of the current anti-diagonal window; the native objects carry the step
counter, `restart()` rewinds it.
"""
import torch
from torch import nn

from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule
from .base import set_weight


class _Restartable(BaseOpModule):
    def restart(self):
        for op in self.op.values():
            op.restart()


class EntropyContextNew(BaseOpModule):

    def __init__(self, npart, rt=18, opt=False, device=0, time_it=False):
        super().__init__(device)
        weight = set_weight(npart, opt)
        self.op = {gid: PCONV.EntropyContextOp(npart, rt, weight, gid, time_it) for gid in self.device_list}

    def setup_context(self, w):
        for op in self.op.values():
            op.start_context(w)

    def get_addr(self, gid):
        return self.op[gid].addr()


class EntropyAdd(_Restartable):
    """y += x on the wavefront cells (residual connection), in place."""

    def __init__(self, npart, channel, ngroup, pad, ctx: EntropyContextNew, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.EntropyAddOp(npart, channel, ngroup, pad, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x, y):
        return self.native(x).forward(x, y)[0]


class EntropyCtxPadRun2(_Restartable):
    """Incremental causal halo / right-wrap fill of the cells that became available this step, in place."""

    def __init__(self, pad, npart, ngroup, ctx: EntropyContextNew, input=False, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.EntropyCtxPadRun2Op(pad, npart, ngroup, input, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(x)[0]


class DExtract2(_Restartable):
    """Gather the channel-group values of the current window; returns (values, cpu int count)."""

    def __init__(self, npart, nchannel, label, ctx: EntropyContextNew, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.DExtract2Op(npart, nchannel, label, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        out = self.native(x).forward(contiguous(x))
        return out[0], out[1]


class DExtract2Batch(_Restartable):
    """Gather the 3 nets' GMM parameters of the current window into [logits | delta | mean] planes."""

    def __init__(self, npart, nchannel, ctx: EntropyContextNew, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.DExtract2Op(npart, nchannel, True, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        out = self.native(x).forward_batch(x)
        return out[0], out[1]


class DInput2(_Restartable):
    """Scatter last step's symbols (+bias) into the padded network input, replicated for the 3 nets."""

    def __init__(self, nchannel, npart, ctx: EntropyContextNew, pad=0, bias=0, repeat=1, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.DInput2Op(nchannel, npart, pad, bias, repeat, ctx.get_addr(gid), gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]


class EntropyConv2(_Restartable):
    """Masked grouped 5x5 convolution on the wavefront cells (single net)."""

    def __init__(self, npart, ngroup, c_in, c_out, kernel_size, ctx: EntropyContextNew, pad_in=2, pad_out=2, hidden=False,
                 act=True, device=0, time_it=False):
        super().__init__(device)
        constrain = 6 if hidden else 5
        channel, nout = ngroup * c_in, ngroup * c_out
        self.op = {gid: PCONV.EntropyConv2Op(npart, channel, ngroup, nout, kernel_size, constrain, pad_in, pad_out,
                                             ctx.get_addr(gid), gid, time_it) for gid in self.device_list}
        self.weight = nn.Parameter(torch.rand((nout, channel, kernel_size, kernel_size), dtype=torch.float32))
        self.bias = nn.Parameter(torch.zeros((nout), dtype=torch.float32))
        self.act = act
        self.relu = nn.Parameter(torch.zeros((nout), dtype=torch.float32)) if act else None

    def forward(self, x):
        op = self.native(x)
        if self.act:
            return op.forward_act(x, self.weight.data, self.bias.data, self.relu.data)[0]
        return op.forward(x, self.weight.data, self.bias.data)[0]


class EntropyConv2Batch(_Restartable):
    """The three nets (mixture logits, delta, mean) evaluated together; batch order = [logits, delta, mean]."""

    def __init__(self, npart, ngroup, c_in, c_out, kernel_size, ctx: EntropyContextNew, pad_in=2, pad_out=2, batch=3,
                 hidden=False, act=True, device=0, time_it=False):
        super().__init__(device)
        constrain = 6 if hidden else 5
        channel, nout = ngroup * c_in, ngroup * c_out
        self.op = {gid: PCONV.EntropyConv2Op(npart, channel, ngroup, nout, kernel_size, constrain, pad_in, pad_out,
                                             ctx.get_addr(gid), gid, time_it) for gid in self.device_list}
        self.weight = nn.Parameter(torch.rand((batch, nout, channel, kernel_size, kernel_size), dtype=torch.float32))
        self.bias = nn.Parameter(torch.rand((batch, nout), dtype=torch.float32))
        self.act = act
        self.relu = nn.Parameter(torch.rand((batch, nout), dtype=torch.float32)) if act else None

    def forward(self, x):
        op = self.native(x)
        if self.act:
            return op.forward_act_batch(x, self.weight.data, self.bias.data, self.relu.data)[0]
        return op.forward_batch(x, self.weight.data, self.bias.data)[0]


class EntropyConvD(nn.Module):

    def __init__(self, ngroups, cin, cout, hidden, npart, out_layer: bool, ctx: EntropyContextNew, device_id, act=True):
        super().__init__()
        pad_out = 0 if out_layer else 2
        self.pad = EntropyCtxPadRun2(2, npart, ngroups, ctx, not hidden, device=device_id)
        self.conv = EntropyConv2(npart, ngroups, cin, cout, 5, ctx, 2, pad_out, hidden=hidden, act=act, device=device_id)

    def forward(self, x):
        return self.conv(self.pad(x))


class EntropyResidualBlockD(nn.Module):

    def __init__(self, ngroups, cpn, npart, ctx: EntropyContextNew, device_id=0):
        super().__init__()
        self.conv1 = EntropyConvD(ngroups, cpn, cpn, True, npart, False, ctx, device_id, True)
        self.conv2 = EntropyConvD(ngroups, cpn, cpn, True, npart, False, ctx, device_id, True)
        self.add = EntropyAdd(npart, cpn * ngroups, ngroups, 2, ctx, device=device_id)

    def forward(self, x):
        return self.add(self.conv2(self.conv1(x)), x)
