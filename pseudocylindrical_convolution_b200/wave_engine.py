"""Native wavefront engine: drives pcx_wave_encode / pcx_wave_decode (include/pcx.h) for an EntEncoder / EntDecoder.

The operator-by-operator loop of pseudo_codec.py:97-114 / :145-160 costs ~35 Python -> ctypes -> launch round trips and two
blocking copies per step; the engine runs the identical launch sequence (same kernels, hence bit-identical CDF tables and
bitstreams - tests/test_gpu_codec.py compares the two paths) from one native call, copies only the live int32 CDF rows to
pinned memory and calls the host range coder from the same loop."""
import ctypes as C

import numpy as np
import torch

from ._lib import WaveNet, call, int_array


def _p(t):
    return t.data_ptr() if t is not None else None


class WaveEngine:
    def __init__(self, ent):
        """ent: pseudo_codec._EntBase (EntEncoder / EntDecoder) - supplies the weights, geometry context and constants."""
        self.ent = ent
        self._key = None

    def _build(self, h, W, nimg, device):
        ent = self.ent
        from . import config
        key = (h, W, nimg, str(device), config.WAVE_CHUNK_ROWS)
        if self._key == key:
            return
        npart, G, pad, nb = ent.npart, ent.ngroup, 2, 3
        ctx = ent.ctx2.op[device.index]
        wl = ctx.widths(h, W)
        band, row, col, tw = ctx.halo(1, h, W, pad)
        items, pstart = ctx.pad_items(h, W, pad)
        d_order, start = ctx.order(h, W)
        Hf = h * npart
        planes = nb * nimg * npart
        z = lambda *shape: torch.zeros(shape, dtype=torch.float32, device=device)
        convs = [ent.net[0].conv] + [c for i in range(1, 6) for c in (ent.net[i].conv1.conv, ent.net[i].conv2.conv)] + [ent.net[6].conv]
        bufs = [z(planes, G, h + 2 * pad, W + 2 * pad)]
        for li, cv in enumerate(convs):
            po = 0 if li == len(convs) - 1 else pad
            bufs.append(z(planes, cv.weight.shape[1], h + 2 * po, W + 2 * po))
        net = WaveNet()
        net.nlayers, net.nb, net.nimg, net.npart, net.G, net.h, net.W, net.pad = len(convs), nb, nimg, npart, G, h, W, pad
        net.nstep, net.ng = 8, 3
        net.gmm_bias, net.gmm_total, net.gmm_beta, net.input_bias = float(ent.bias), 65536.0, 1e-6, -float(ent.bias)
        self._wl = int_array(wl)
        self._pstart = np.ascontiguousarray(pstart, np.int32)
        self._start = np.ascontiguousarray(start, np.int32)
        net.wl = C.cast(self._wl, C.c_void_p)
        net.d_band, net.d_row, net.d_col, net.d_tw = _p(band), _p(row), _p(col), _p(tw)
        net.d_items, net.h_pstart = _p(items), self._pstart.ctypes.data
        net.d_order, net.h_start = _p(d_order), self._start.ctypes.data
        go_last = convs[-1].weight.shape[1] // G
        self.params = z(nb, go_last, Hf * nimg, W)
        # row capacity: at least one wavefront step (stepwise engine); the one-shot encoder emits chunks of up to WAVE_CHUNK_ROWS rows
        # per image so that host coding of one chunk overlaps the device work and the copy of the next
        ncell = h * sum(wl)
        rows = nimg * max(min(Hf, W) * G + 16, min(ncell * G, config.WAVE_CHUNK_ROWS))
        self.cdf = torch.zeros((rows, 9), dtype=torch.int32, device=device)
        self.lab = torch.zeros((rows,), dtype=torch.int32, device=device)
        self.steptab = torch.zeros((2 * (Hf + W + G),), dtype=torch.int32, device=device)
        self.prev = z(nimg, Hf * W)
        net.d_params, net.d_cdf, net.d_prev = _p(self.params), _p(self.cdf), _p(self.prev)
        net.cdf_rows, net.d_lab, net.d_steptab = rows, _p(self.lab), _p(self.steptab)
        for li, cv in enumerate(convs):
            l = net.layers[li]
            l.weight, l.bias = _p(cv.weight.data), _p(cv.bias.data)
            l.act = _p(cv.relu.data) if cv.act else None
            l.in_, l.out = _p(bufs[li]), _p(bufs[li + 1])
            l.gi, l.go = cv.weight.shape[2] // G, cv.weight.shape[1] // G
            l.pad_out = 0 if li == len(convs) - 1 else pad
            l.constrain = 5 if li == 0 else 6
            l.input_layer = 1 if li == 0 else 0
            # residual blocks: the second conv of block i adds the block's input (= output buffer of layer 2i)
            l.add = _p(bufs[li - 1]) if (li >= 2 and li % 2 == 0 and li < len(convs) - 1) else None
        self.net, self.bufs, self._keep = net, bufs, (band, row, col, tw, items, d_order, convs)
        self._key = key

    @staticmethod
    def _handles(coders):
        coders = coders if isinstance(coders, (list, tuple)) else [coders]
        return (C.c_void_p * len(coders))(*[c._h for c in coders]), len(coders)

    def encode(self, data, coders):
        """data: (nimg*npart, G, h, W) symbols (already PseudoFill'ed); coders: one started coder.coder per image."""
        NN, G, h, W = data.shape
        nimg = NN // self.ent.npart
        hs, nc = self._handles(coders)
        assert nc == nimg, "one coder per image"
        with torch.cuda.device(data.device):
            self._build(h, W, nimg, data.device)
            n = C.c_longlong(0)
            from . import config
            fn = "pcx_wave_encode_full" if config.WAVE_ENCODE_FULL else "pcx_wave_encode"
            call(fn, C.byref(self.net), C.c_void_p(data.data_ptr()), hs, C.byref(n),
                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        return n.value

    def decode(self, h, W, coders, device):
        """returns the decoded symbol tensor (nimg*npart, G, h, W) as float; coders: one started decoder per image"""
        hs, nimg = self._handles(coders)
        with torch.cuda.device(device):
            self._build(h, W, nimg, device)
            n = C.c_longlong(0)
            call("pcx_wave_decode", C.byref(self.net), hs, C.byref(n), C.c_void_p(torch.cuda.current_stream().cuda_stream))
            b = self.bufs[0]
            return (b[:nimg * self.ent.npart, :, 2:-2, 2:-2] + self.ent.bias).contiguous()
