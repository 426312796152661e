"""Analysis / synthesis transforms of the 360-degree codec (reference: model_zoo_v2.py:8-211).

Same class names, constructor signatures and state_dict keys as the reference, so its
checkpoints load strictly.  The execution differs: the reference runs PseudoPadV2 -> nn.Conv2d (cuDNN) ->
PReLU/Sigmoid/add -> PseudoFillV2 as separate passes; here every convolution goes through libpcx's
`pcx_conv2d_fwd` with bias, activation, gate/residual and the invalid-column fill fused into its epilogue.
The nn.Conv2d / nn.PReLU members are parameter containers only - their forward() is never called, so no
cuDNN kernel runs on this path.
"""
import ctypes as C

import torch
from torch import nn

from . import config
from ._lib import ConvDesc, call
from .PCONV_operator import Dtow, PseudoContextV2, PseudoFillV2, PseudoGDNV2, PseudoPadV2

ACT_NONE, ACT_PRELU, ACT_SIGMOID = 0, 1, 2


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def pconv(x, conv: nn.Conv2d, npart, wl_out, prelu: nn.PReLU = None, sigmoid=False, mul=None, residual=None, impl=None):
    """y = fill(residual + mul * act(conv(x) + bias)) on an already padded tile tensor x (NN, Ci, Hi, Wi).

    wl_out[g] = number of valid output columns of band g (columns beyond are written as zero)."""
    assert x.is_contiguous() and x.dtype == torch.float32
    NN, Ci, Hi, Wi = x.shape
    k, s = conv.kernel_size[0], conv.stride[0]
    Co = conv.out_channels
    Ho, Wo = (Hi - k) // s + 1, (Wi - k) // s + 1
    y = torch.empty((NN, Co, Ho, Wo), dtype=x.dtype, device=x.device)
    d = ConvDesc()
    d.N, d.npart = NN // npart, npart
    d.Ci, d.Hi, d.in_pitch = Ci, Hi, Wi
    d.Co, d.Ho, d.Wo = Co, Ho, Wo
    d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = Ho, Wo, 0, 0
    d.k, d.stride = k, s
    d.act = ACT_PRELU if prelu is not None else (ACT_SIGMOID if sigmoid else ACT_NONE)
    d.impl = 1 if impl is None else impl          # NCHW tensors: the fp32 direct form (the tensor-core path is transforms_nhwc)
    d.aux_rows, d.aux_pitch, d.aux_y0, d.aux_x0 = Ho, Wo, 0, 0
    for g in range(npart):
        d.wl_out[g] = min(int(wl_out[g]), Wo)
    for aux in (mul, residual):
        if aux is not None:
            assert aux.is_contiguous() and tuple(aux.shape) == (NN, Co, Ho, Wo)
    with torch.cuda.device(x.device):
        call("pcx_conv2d_fwd", C.byref(d), _ptr(x), _ptr(conv.weight.data), _ptr(conv.bias.data if conv.bias is not None else None),
             _ptr(prelu.weight.data if prelu is not None else None), _ptr(mul), _ptr(residual), _ptr(y),
             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    return y


def _widths(ctx: PseudoContextV2, x, h, W):
    return ctx.op[x.device.index].widths(h, W)


class ClipData(nn.Module):
    """Leaky clip to [0,1] with slope 0.01 outside (reference: model_zoo_v2.py:8-34)."""

    def forward(self, x):
        return torch.where(x < 0, x * 0.01, torch.where(x > 1, 1 + (x - 1) * 0.01, x))


def _refresh_hook(module, incompatible_keys):
    module.refresh()               # (a post hook must return None)


class _Block(nn.Module):
    def __init__(self, npart, ctx):
        super().__init__()
        self.npart = npart
        self._ctx = [ctx]          # shared geometry object, deliberately not registered as a submodule
        self._nhwc = [None]
        self.register_load_state_dict_post_hook(_refresh_hook)

    def wl(self, x, h, W):
        return _widths(self._ctx[0], x, h, W)

    def refresh(self):
        """Drop everything derived from the parameters: packed TF32 weights, GDN beta' / gamma', the captured CUDA graph.  The
        caches key on Tensor._version, which in-place updates through `.data` (a common idiom, also in checkpoint surgery) do
        not bump: call this after such an edit.  load_state_dict() and .to() / .cuda() call it through the hooks below."""
        self.__dict__.pop("_pcx_graph", None)
        for m in self.modules():
            if isinstance(m, _Block) and m._nhwc[0] is not None:
                m._nhwc[0]._packed.clear()
            if hasattr(m, "_eff"):
                m._eff = None
        return self

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.refresh()
        return out

    def _runner(self):
        """channels-last / tensor-core executor of this transform (transforms_nhwc.Runner), created on first use"""
        if self._nhwc[0] is None:
            from .transforms_nhwc import Runner
            gid = next(self.parameters()).device.index
            self._nhwc[0] = Runner(self._ctx[0].op[gid], self.npart)
        return self._nhwc[0]


class ResidualBlock(_Block):
    """x + conv1x1(PReLU(conv3x3(PReLU(conv1x1(pad1(x))))))  - reference :36-53"""

    def __init__(self, channels, npart, ctx: PseudoContextV2, device_id=0):
        super().__init__(npart, ctx)
        self.pad = PseudoPadV2(1, npart, ctx, device=device_id)
        self.conv1 = nn.Conv2d(channels, channels // 2, 1, 1)
        self.relu1 = nn.PReLU(channels // 2)
        self.conv2 = nn.Conv2d(channels // 2, channels // 2, 3, 1)
        self.relu2 = nn.PReLU(channels // 2)
        self.conv3 = nn.Conv2d(channels // 2, channels, 1, 1)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        wl = self.wl(x, h, W)
        tx = self.pad(x)
        y = pconv(tx, self.conv1, self.npart, [w + 2 for w in wl], prelu=self.relu1)
        y = pconv(y, self.conv2, self.npart, wl, prelu=self.relu2)
        return pconv(y, self.conv3, self.npart, wl, residual=x)


class AttentionBlock(_Block):
    """x + trunk(x) * sigmoid(conv1x1(attention(x)))  - reference :55-76"""

    def __init__(self, channels, npart, ctx: PseudoContextV2, device_id=0):
        super().__init__(npart, ctx)
        self.trunk = nn.Sequential(*[ResidualBlock(channels, npart, ctx, device_id) for _ in range(3)])
        self.attention = nn.Sequential(*[ResidualBlock(channels, npart, ctx, device_id) for _ in range(3)],
                                       nn.Conv2d(channels, channels, 1, 1, 0), nn.Sigmoid())
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        wl = self.wl(x, h, W)
        t = self.trunk(x)
        a = x
        for i in range(3):
            a = self.attention[i](a)
        return pconv(a, self.attention[3], self.npart, wl, sigmoid=True, mul=t, residual=x)


class ResidualBlockV2(_Block):
    """x + PReLU(conv3x3(PReLU(conv3x3(pad2(x)))))  - reference :78-93"""

    def __init__(self, channels, npart, ctx: PseudoContextV2, device_id):
        super().__init__(npart, ctx)
        self.pad = PseudoPadV2(2, npart, ctx, device=device_id)
        self.conv1 = nn.Conv2d(channels, channels, 3, 1)
        self.relu1 = nn.PReLU(channels)
        self.conv2 = nn.Conv2d(channels, channels, 3, 1)
        self.relu2 = nn.PReLU(channels)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        wl = self.wl(x, h, W)
        tx = self.pad(x)
        y = pconv(tx, self.conv1, self.npart, [w + 2 for w in wl], prelu=self.relu1)
        return pconv(y, self.conv2, self.npart, wl, prelu=self.relu2, residual=x)


class ResidualBlockDown(_Block):
    """conv1x1/2(x) + GDN(conv3x3(pad1(PReLU(conv3x3/2(pad1(x))))))  - reference :95-114"""

    def __init__(self, channels, channel_in, npart, ctx: PseudoContextV2, device_id):
        super().__init__(npart, ctx)
        self.pad1 = PseudoPadV2(1, npart, ctx, device=device_id)
        self.conv1 = nn.Conv2d(channel_in, channels, 3, 2)
        self.relu1 = nn.PReLU(channels)
        self.pad2 = PseudoPadV2(1, npart, ctx, device=device_id)
        self.conv2 = nn.Conv2d(channels, channels, 3, 1)
        self.relu2 = PseudoGDNV2(channels, npart, ctx, device_id)
        self.short_cut = nn.Conv2d(channel_in, channels, 1, 2)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        wl2 = self.wl(x, h // 2, W // 2)
        t = pconv(x, self.short_cut, self.npart, wl2)
        y = pconv(self.pad1(x), self.conv1, self.npart, wl2, prelu=self.relu1)
        y = pconv(self.pad2(y), self.conv2, self.npart, wl2)
        return self.relu2(y, residual=t)


class SphereConv2(_Block):
    """fill(conv3x3/2(pad1(x)))  - reference :116-126"""

    def __init__(self, channel_in, channel_out, npart, ctx: PseudoContextV2, device_id=0):
        super().__init__(npart, ctx)
        self.conv = nn.Conv2d(channel_in, channel_out, 3, 2, 0)
        self.pad = PseudoPadV2(1, npart, ctx, device=device_id)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        return pconv(self.pad(x), self.conv, self.npart, self.wl(x, h // 2, W // 2))


class EncoderV2(_Block):
    """Analysis transform 3 -> 192 channels, /16 resolution, sigmoid code  - reference :129-151"""

    def __init__(self, channels, code_channels, npart, ctx: PseudoContextV2, device_id):
        super().__init__(npart, ctx)
        self.net = nn.Sequential(
            ResidualBlockDown(channels, 3, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            ResidualBlockDown(channels, channels, npart, ctx, device_id),
            AttentionBlock(channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            ResidualBlockDown(channels, channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            SphereConv2(channels, channels, npart, ctx, device_id),
            AttentionBlock(channels, npart, ctx, device_id),
            nn.Conv2d(channels, code_channels, 1, 1),
        )
        self.act = nn.Sigmoid()
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        for i in range(9):
            x = self.net[i](x)
        h, W = x.shape[2:]
        return pconv(x, self.net[9], self.npart, self.wl(x, h, W), sigmoid=True)


class ResidualBlockUp(_Block):
    """d2w(conv1x1(x)) + IGDN(conv3x3(pad1(d2w(PReLU(conv3x3(pad1(x)))))))  - reference :153-175"""

    def __init__(self, channels, npart, ctx: PseudoContextV2, device_id):
        super().__init__(npart, ctx)
        self.pad1 = PseudoPadV2(1, npart, ctx, device=device_id)
        self.conv1 = nn.Conv2d(channels, channels * 4, 3, 1)
        self.relu1 = nn.PReLU(channels * 4)
        self.dtow1 = Dtow(2, True, device_id)
        self.pad2 = PseudoPadV2(1, npart, ctx, device=device_id)
        self.conv2 = nn.Conv2d(channels, channels, 3, 1)
        self.relu2 = PseudoGDNV2(channels, npart, ctx, device_id, inverse=True)
        self.short_cut = nn.Conv2d(channels, channels * 4, 1, 1)
        self.dtow2 = Dtow(2, True, device_id)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        wl = self.wl(x, h, W)
        wl2 = self.wl(x, 2 * h, 2 * W)
        # The reference does not trim conv1 / short_cut before Dtow (:166-173), and round-half-up band widths do not always double
        # (code width 104 -> 208: band 0 has 24 -> 49 columns): the last valid up-sampled column then comes from conv column
        # wl[g], one beyond the band.  Keep every column the up-sampled band needs.
        wl_up = [min(W, max(int(a), (int(b) + 1) // 2)) for a, b in zip(wl, wl2)]
        br1 = pconv(self.pad1(x), self.conv1, self.npart, wl_up, prelu=self.relu1)
        br1 = self.dtow1(br1)
        br1 = pconv(self.pad2(br1), self.conv2, self.npart, wl2)
        br2 = self.dtow2(pconv(x, self.short_cut, self.npart, wl_up))
        return self.relu2(br1, residual=br2)


class SphereConvOld(_Block):
    """fill(conv1x1(x))  - reference :177-186"""

    def __init__(self, npart, channel_in, channel_out, ctx: PseudoContextV2, device_id=0):
        super().__init__(npart, ctx)
        self.conv = nn.Conv2d(channel_in, channel_out, 1, 1)
        self.trim = PseudoFillV2(0, npart, ctx, device=device_id)

    def forward(self, x):
        h, W = x.shape[2:]
        return pconv(x, self.conv, self.npart, self.wl(x, h, W))


class DecoderV2(_Block):
    """Synthesis transform 192 -> 3 channels, x16 resolution  - reference :189-211"""

    def __init__(self, channels, code_channels, npart, ctx: PseudoContextV2, device_id):
        super().__init__(npart, ctx)
        self.net = nn.Sequential(
            SphereConvOld(npart, code_channels, channels, ctx, device_id),
            AttentionBlock(channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            ResidualBlockUp(channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            ResidualBlockUp(channels, npart, ctx, device_id),
            AttentionBlock(channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            ResidualBlockUp(channels, npart, ctx, device_id),
            ResidualBlockV2(channels, npart, ctx, device_id),
            PseudoPadV2(1, npart, ctx, device=device_id),
            nn.Conv2d(channels, 12, 3, 1),
            Dtow(2, True, device_id),
        )

    def forward(self, x):
        for i in range(10):
            x = self.net[i](x)
        h, W = x.shape[2:]
        wl, wl2 = self.wl(x, h, W), self.wl(x, 2 * h, 2 * W)
        wl_up = [min(W, max(int(a), (int(b) + 1) // 2)) for a, b in zip(wl, wl2)]      # untrimmed before Dtow, as in ResidualBlockUp
        y = pconv(self.net[10](x), self.net[11], self.npart, wl_up)
        return self.net[12](y)

    def widths_double(self, x_like, h, W):
        """True when, at every up-sampling stage of this transform, no band is wider than twice its width one scale below.
        The channels-last path keeps its wrap columns inside the activation buffers and therefore cannot reproduce the
        reference's untrimmed column beyond the band; such image widths (not multiples of 1024, e.g. 1664) run the exact
        NCHW path (PseudoDecoder.reconstruct)."""
        for _ in range(4):
            a, b = self.wl(x_like, h, W), self.wl(x_like, 2 * h, 2 * W)
            if any(int(q) > 2 * int(p) for p, q in zip(a, b)):
                return False
            h, W = 2 * h, 2 * W
        return True
