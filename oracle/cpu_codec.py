"""CPU port of the reference codec (TEST INFRASTRUCTURE / CPU BASELINE ONLY - same import rules as oracle.py).

The reference has no CPU path (every operator is a CUDA kernel), so "the reference's CPU implementation"
of pseudo_codec.py --enc / --dec is this transliteration: the reference's layer graph (model_zoo_v2.py:36-211) and codec
loops (pseudo_codec.py:97-114, :145-160, :178-213) evaluated with

  * oracle/pcx_oracle.c for every custom operator (slice, uslice, pad, fill, quant, dquant, dtow, the wavefront ops, the
    GMM table) - pinned bit for bit to outputs of the unmodified reference extension (tests/test_golden_oracle.py);
  * torch-CPU fp32 `conv2d` / `prelu` / `sigmoid` where the reference calls ATen / cuDNN;
  * the reference's own arithmetic coder, compiled from /root/reference/coder (oracle/_ref/coder_ref.so), when present; the
    product's byte-identical host coder otherwise (stated in `coder_kind`).

It serves two purposes: (1) `bench.py --impl reference` / `cpu_baseline` time it on the host cores; (2) it is the INDEPENDENT
checker of the product's transform topology - in particular of the synthesis side (DecoderV2 / ResidualBlockUp / IGDN wiring),
which no other oracle covers (tests/test_gpu_codec.py::test_synthesis_transform_vs_cpu_port).
State dicts use the reference's keys, so the same checkpoints feed both implementations.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch
import torch.nn.functional as F

from . import oracle as orc

NPART = 16


def _np(t):
    return np.ascontiguousarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float32)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


class CpuCodec:
    """PseudoEncoder + PseudoDecoder of pseudo_codec.py:164-213 on the host.  sd: merged state dict with the reference's keys
    (`encoder.*`, `decoder.*`, `quant.weight`, `ent.*`)."""

    def __init__(self, sd, valid_dim=56, threads=None, coder_mod=None, coder_kind="port"):
        self.sd = {k: v.detach().cpu().float() for k, v in sd.items()}
        self.vd = int(valid_dim)
        self.G = self.vd // 4
        self.threads = int(threads or os.cpu_count() or 1)
        self.coder_mod, self.coder_kind = coder_mod, coder_kind
        self.pool = ThreadPoolExecutor(self.threads)
        orc.build()
        self._geom = {}

    # ------------------------------------------------------------------------------------------ helpers
    def wl(self, h, W):
        return orc.band_widths(orc.W64_NPART16, h * NPART, W)

    def _par(self, fn, x):
        """custom operators act per (plane, channel): split the channel axis over the host threads (ctypes drops the GIL)"""
        Cc = x.shape[1]
        n = min(self.threads, Cc)
        if n <= 1:
            return fn(x)
        step = -(-Cc // n)
        parts = list(self.pool.map(lambda c: fn(np.ascontiguousarray(x[:, c:c + step])), range(0, Cc, step)))
        return np.concatenate(parts, axis=1)

    def pad(self, x, p):                      # PseudoPadV2 (pseudo_pad.cu:39-96)
        wl = self.wl(x.shape[2], x.shape[3])
        return self._par(lambda a: orc.pseudo_pad(a, wl, p), _np(x))

    def fill(self, x):                        # PseudoFillV2(0) (pseudo_fill_cuda.cu:28-43)
        wl = self.wl(x.shape[2], x.shape[3])
        return self._par(lambda a: orc.pseudo_fill(a, wl), _np(x))

    def conv(self, x, name, stride=1):
        w, b = self.sd[name + ".weight"], self.sd.get(name + ".bias")
        return F.conv2d(_t(x), w, b, stride=stride)

    def prelu(self, y, name):
        return F.prelu(y, self.sd[name + ".weight"])

    # ------------------------------------------------------------------------------------------ blocks (model_zoo_v2.py)
    def residual_block(self, p, x):           # :36-53
        y = self.prelu(self.conv(self.pad(x, 1), p + ".conv1"), p + ".relu1")
        y = self.prelu(F.conv2d(y, self.sd[p + ".conv2.weight"], self.sd[p + ".conv2.bias"]), p + ".relu2")
        y = F.conv2d(y, self.sd[p + ".conv3.weight"], self.sd[p + ".conv3.bias"])
        return self.fill(_t(x) + y)

    def attention_block(self, p, x):          # :55-76
        t = a = x
        for i in range(3):
            t = self.residual_block("%s.trunk.%d" % (p, i), t)
            a = self.residual_block("%s.attention.%d" % (p, i), a)
        a = torch.sigmoid(self.conv(a, p + ".attention.3"))
        return self.fill(_t(x) + _t(t) * a)

    def residual_block_v2(self, p, x):        # :78-93
        y = self.prelu(self.conv(self.pad(x, 2), p + ".conv1"), p + ".relu1")
        y = self.prelu(F.conv2d(y, self.sd[p + ".conv2.weight"], self.sd[p + ".conv2.bias"]), p + ".relu2")
        return self.fill(_t(x) + y)

    def gdn(self, p, x, inverse):             # PseudoContextV2.py:186-216, GDN.py:6-22
        x = _t(x)
        ch = x.shape[1]
        ro = torch.tensor([2.0 ** -18], dtype=torch.float32)
        pedestal = ro ** 2
        beta_bound = (1e-6 + ro ** 2) ** .5
        mask = _t(self.fill(np.ones(tuple(x.shape), np.float32)))
        x = x * mask
        beta = torch.max(self.sd[p + ".beta"], torch.ones(ch) * beta_bound) ** 2 - pedestal
        gamma = torch.max(self.sd[p + ".gamma"], torch.ones(ch, ch) * ro) ** 2 - pedestal
        norm = torch.sqrt(F.conv2d(x ** 2, gamma.view(ch, ch, 1, 1), beta))
        norm = norm * mask + 1 - mask
        return x * norm if inverse else x / norm

    def residual_block_down(self, p, x):      # :95-114
        t = self.conv(x, p + ".short_cut", stride=2)
        y = self.prelu(self.conv(self.pad(x, 1), p + ".conv1", stride=2), p + ".relu1")
        y = self.gdn(p + ".relu2", self.conv(self.pad(y, 1), p + ".conv2"), False)
        return self.fill(t + y)

    def residual_block_up(self, p, x):        # :153-175
        b1 = self.prelu(self.conv(self.pad(x, 1), p + ".conv1"), p + ".relu1")
        b1 = orc.dtow(_np(b1), 2, True)
        b1 = self.gdn(p + ".relu2", self.conv(self.pad(b1, 1), p + ".conv2"), True)
        b2 = _t(orc.dtow(_np(self.conv(x, p + ".short_cut")), 2, True))
        return self.fill(b1 + b2)

    # ------------------------------------------------------------------------------------------ transforms
    def analysis(self, erp):
        """EncoderV2 after SphereSlice (:129-151; pseudo_codec.py:178-181): ERP (N,3,H,W) -> code (16N,192,H/256,W/16) in [0,1]."""
        erp = _np(erp)
        H, W = erp.shape[2:]
        x = orc.sphere_slice(erp, self.wl(H // NPART, W))
        e = "encoder.net."
        x = self.residual_block_down(e + "0", x)
        x = self.residual_block_v2(e + "1", x)
        x = self.residual_block_down(e + "2", x)
        x = self.attention_block(e + "3", x)
        x = self.residual_block_v2(e + "4", x)
        x = self.residual_block_down(e + "5", x)
        x = self.residual_block_v2(e + "6", x)
        x = self.fill(self.conv(self.pad(x, 1), e + "7.conv", stride=2))            # SphereConv2 :116-126
        x = self.attention_block(e + "8", x)
        return self.fill(torch.sigmoid(self.conv(x, e + "9")))

    def synthesis(self, code_f):
        """DecoderV2 + SphereUslice + ClipData (:189-211, :8-34; pseudo_codec.py:211-213): (16N,192,h,w) -> ERP (N,3,256h,16w)."""
        d = "decoder.net."
        x = self.fill(self.conv(code_f, d + "0.conv"))                                # SphereConvOld :177-186
        x = self.attention_block(d + "1", x)
        x = self.residual_block_v2(d + "2", x)
        x = self.residual_block_up(d + "3", x)
        x = self.residual_block_v2(d + "4", x)
        x = self.residual_block_up(d + "5", x)
        x = self.attention_block(d + "6", x)
        x = self.residual_block_v2(d + "7", x)
        x = self.residual_block_up(d + "8", x)
        x = self.residual_block_v2(d + "9", x)
        x = self.conv(self.pad(x, 1), d + "11")
        x = orc.dtow(_np(x), 2, True)
        erp = orc.sphere_uslice(x, self.wl(x.shape[2], x.shape[3]))
        y = erp.copy()                                                                # ClipData
        y[erp < 0] = erp[erp < 0] * 0.01
        y[erp > 1] = 1 + (erp[erp > 1] - 1) * 0.01
        return y

    def symbols(self, erp):
        """pseudo_codec.py:178-184: quantise, keep valid_dim channels, depth-to-width -> (16N, G, H/128, W/8) symbols 0..7"""
        lat = self.analysis(erp)
        steps = orc.quant_steps(_np(self.sd["quant.weight"]))
        _, sym, _ = orc.pseudo_quant(lat, steps, self.wl(lat.shape[2], lat.shape[3]))
        return orc.dtow(np.ascontiguousarray(sym[:, :self.vd]), 2, True)

    def reconstruct(self, hcode):
        """pseudo_codec.py:207-213"""
        code_i = orc.dtow(_np(hcode), 2, False)
        centres = orc.dquant_centres(_np(self.sd["quant.weight"]))
        n, _, h, w = code_i.shape
        code = np.zeros((n, 192, h, w), np.float32)
        code[:, :self.vd] = orc.pseudo_dquant(code_i, centres[:self.vd], self.wl(h, w))
        return self.synthesis(code)

    # ------------------------------------------------------------------------------------------ context model
    def _ctx(self, h, W):
        key = (h, W)
        if key not in self._geom:
            self._geom[key] = orc.CtxGeom(self.wl(h, W), h, W, 2)
        return self._geom[key]

    def _layers(self):
        e = "ent.net."
        names = [e + "0.conv"] + [e + "%d.%s.conv" % (i, c) for i in range(1, 6) for c in ("conv1", "conv2")] + [e + "6.conv"]
        return [(_np(self.sd[n + ".weight"]), _np(self.sd[n + ".bias"]), _np(self.sd[n + ".relu"]) if (n + ".relu") in self.sd else None)
                for n in names]

    def _wave(self, h, W, step_fn):
        """The wavefront loop shared by EntEncoder / EntDecoder (pseudo_codec.py:97-114, :145-160) for ONE image.
        step_fn(s, cdf (n,9) int32 or None, n) -> symbols (n,) float32 of that step (the coder decides which way they flow)."""
        G, geom, Hf = self.G, self._ctx(h, W), h * NPART
        layers = self._layers()
        NN = 3 * NPART
        b_in = np.zeros((NN, G, h + 4, W + 4), np.float32)
        outs = [np.zeros((NN, 3 * G, h + 4, W + 4), np.float32) for _ in range(11)] + [np.zeros((NN, 3 * G, h, W), np.float32)]
        o_ext = np.zeros((3, 3, Hf, W), np.float32)
        prev = np.zeros(Hf * W, np.float32)

        def layer(k, x, s):
            w, b, a = layers[k]
            orc.ctx_pad_step(x, geom, G, s - 1 if k == 0 else s)
            orc.ctx_conv_step(x, w, b, a, outs[k], geom, G, 1, 2, 0 if k == 11 else 2, 5 if k == 0 else 6, s, self.pool, self.threads)
            return outs[k]

        for s in range(geom.nsteps(G)):
            orc.dinput_step(prev, b_in, geom, G, 1, 2, -3.5, 3, s)
            x = layer(0, b_in, s)
            for blk in range(5):
                y = layer(2 + 2 * blk, layer(1 + 2 * blk, x, s), s)
                x = orc.ctx_add_step(y, x, geom, G, 2, s)
            n = orc.dextract_step(layer(11, x, s), o_ext, geom, G, s, True)
            cdf = None
            if n > 0:
                z = o_ext.reshape(3, -1)[:, :n * 3]
                cdf, _, _ = orc.gmm_table(z[0].reshape(n, 3), z[1].reshape(n, 3), z[2].reshape(n, 3))
            sym = step_fn(s, cdf, n)
            prev = np.zeros(Hf * W, np.float32)
            prev[:n] = sym
        orc.dinput_step(prev, b_in, geom, G, 1, 2, -3.5, 3, geom.nsteps(G))
        return b_in

    def _coder(self, path):
        if self.coder_mod is not None:
            return self.coder_mod.coder(path)
        from pseudocylindrical_convolution_b200 import coder as mine      # byte-identical (tests/test_oracle_cpu.py)
        return mine.coder(path)

    def entropy_encode(self, hcode, path):
        """EntEncoder.forward (pseudo_codec.py:97-114) for one image: hcode (16, G, h, W)"""
        hcode = _np(hcode)
        _, G, h, W = hcode.shape
        data = orc.pseudo_fill(hcode, self.wl(h, W))
        geom = self._ctx(h, W)
        o_lab = np.zeros((1, 1, h * NPART, W), np.float32)
        c = self._coder(path)
        c.start_encoder()

        def step(s, cdf, n):
            m = orc.dextract_step(data, o_lab, geom, G, s, False)
            lab = o_lab.reshape(-1)[:m].copy()
            if n > 0:
                c.encodes(torch.from_numpy(cdf), 8, torch.from_numpy(lab[:n].astype(np.int32)), n)
            return lab[:n]

        self._wave(h, W, step)
        c.end_encoder()
        return data

    def entropy_decode(self, path, h, W):
        """EntDecoder.forward (pseudo_codec.py:145-160) for one image -> (16, G, h, W) symbols"""
        c = self._coder(path)
        c.start_decoder()

        def step(s, cdf, n):
            if n == 0:
                return np.zeros(0, np.float32)
            out = c.decodes(torch.from_numpy(cdf), 8, n)
            return _np(out)[:n]

        b = self._wave(h, W, step)
        code = b[:NPART, :, 2:-2, 2:-2] + 3.5
        return orc.pseudo_fill(np.ascontiguousarray(code), self.wl(h, W))

    # ------------------------------------------------------------------------------------------ pseudo_codec.py entry points
    def encode(self, erp, path):
        """PseudoEncoder.forward (pseudo_codec.py:178-186) for one image (1,3,H,W) in [0,1]"""
        with torch.no_grad():
            torch.set_num_threads(self.threads)
            hcode = self.symbols(erp)
            return self.entropy_encode(hcode, path)

    def decode(self, path, H=512, W=1024):
        """PseudoDecoder.forward (pseudo_codec.py:203-213)"""
        with torch.no_grad():
            torch.set_num_threads(self.threads)
            return self.reconstruct(self.entropy_decode(path, H // 128, W // 8))


def load_reference_coder():
    """oracle/_ref/coder_ref.so = the reference's coder/ compiled where it lies (oracle/build_ref.py); None if absent."""
    import importlib.util
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "coder_ref.so")
    if not os.path.exists(path):
        return None
    try:
        spec = importlib.util.spec_from_file_location("coder_ref", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------------- checkpoints
def synth_state_dicts(valid_dim=56, seed=0, channels=192):
    """Seeded random-init parameters with the reference's key sets, built on the CPU with plain torch modules:
    (encoder file, decoder file, ent file) dicts = what `{prex}_encoder.pt`, `{prex}_decoder.pt`, `{prex}_ent.pt` hold
    (pseudo_codec.py:223-227).  Transforms: torch default inits; GDN and quantiser as the reference constructs them
    (PseudoContextV2.py:159-175, :245-249) with the centres spread over the sigmoid range; entropy net drawn like the training
    net (kaiming-normal over the causal half of the taps, zero bias, delta-net output bias 2, PReLU slope 0.25 - SURVEY.md H6)."""
    import math
    from torch import nn
    ch, G = channels, valid_dim // 4
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, ci, co, k):
        m = nn.Conv2d(ci, co, k)
        with torch.no_grad():
            bound = 1.0 / math.sqrt(ci * k * k)
            m.weight.copy_((torch.rand(m.weight.shape, generator=g) * 2 - 1) * bound)      # kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(fan_in), .)
            m.bias.copy_((torch.rand(m.bias.shape, generator=g) * 2 - 1) * bound)
        sd[name + ".weight"], sd[name + ".bias"] = m.weight.detach().clone(), m.bias.detach().clone()

    def prelu(name, c):
        sd[name + ".weight"] = torch.full((c,), 0.25)

    def gdn(name, c):
        ped = torch.tensor([2.0 ** -18]) ** 2
        sd[name + ".beta"] = torch.sqrt(torch.ones(c) + ped)
        sd[name + ".gamma"] = torch.sqrt(0.1 * torch.eye(c) + ped)

    def rb(p):
        conv(p + ".conv1", ch, ch // 2, 1); prelu(p + ".relu1", ch // 2)
        conv(p + ".conv2", ch // 2, ch // 2, 3); prelu(p + ".relu2", ch // 2)
        conv(p + ".conv3", ch // 2, ch, 1)

    def att(p):
        for i in range(3):
            rb("%s.trunk.%d" % (p, i))
        for i in range(3):
            rb("%s.attention.%d" % (p, i))
        conv(p + ".attention.3", ch, ch, 1)

    def rbv2(p):
        conv(p + ".conv1", ch, ch, 3); prelu(p + ".relu1", ch)
        conv(p + ".conv2", ch, ch, 3); prelu(p + ".relu2", ch)

    def down(p, cin):
        conv(p + ".conv1", cin, ch, 3); prelu(p + ".relu1", ch)
        conv(p + ".conv2", ch, ch, 3); gdn(p + ".relu2", ch)
        conv(p + ".short_cut", cin, ch, 1)

    def up(p):
        conv(p + ".conv1", ch, 4 * ch, 3); prelu(p + ".relu1", 4 * ch)
        conv(p + ".conv2", ch, ch, 3); gdn(p + ".relu2", ch)
        conv(p + ".short_cut", ch, 4 * ch, 1)

    e = "encoder.net."
    down(e + "0", 3); rbv2(e + "1"); down(e + "2", ch); att(e + "3"); rbv2(e + "4"); down(e + "5", ch); rbv2(e + "6")
    conv(e + "7.conv", ch, ch, 3); att(e + "8"); conv(e + "9", ch, ch, 1)
    d = "decoder.net."
    conv(d + "0.conv", ch, ch, 1); att(d + "1"); rbv2(d + "2"); up(d + "3"); rbv2(d + "4"); up(d + "5"); att(d + "6"); rbv2(d + "7")
    up(d + "8"); rbv2(d + "9"); conv(d + "11", ch, 12, 3)
    qw = torch.zeros(ch, 8)
    qw[:, 0] = 0.06
    qw[:, 1:] = math.log(0.125)
    sd["quant.weight"] = qw
    names = ["ent.net.0.conv"] + ["ent.net.%d.%s.conv" % (i, c) for i in range(1, 6) for c in ("conv1", "conv2")] + ["ent.net.6.conv"]
    for li, n in enumerate(names):
        cin = G if li == 0 else 3 * G
        w = torch.randn((3, 3 * G, cin, 5, 5), generator=g) * math.sqrt(2.0 / (cin * 25 / 2.0))
        b = torch.zeros(3, 3 * G)
        if li == len(names) - 1:
            b[1] = 2.0
        sd[n + ".weight"], sd[n + ".bias"] = w, b
        if li < len(names) - 1:
            sd[n + ".relu"] = torch.full((3, 3 * G), 0.25)
    enc = {k: v for k, v in sd.items() if k.startswith("encoder.")}
    enc["quant.weight"] = qw.clone()
    enc["quant.count"] = torch.zeros(ch, 8)
    dec = {k: v for k, v in sd.items() if k.startswith("decoder.")}
    dec["quant.weight"] = qw.clone()
    ent = {k: v for k, v in sd.items() if k.startswith("ent.")}
    return enc, dec, ent


def save_checkpoints(model_dir, prex, valid_dim=56, seed=0):
    """writes {prex}_encoder.pt / _decoder.pt / _ent.pt (the reference's file convention) and returns (paths, merged dict)"""
    os.makedirs(model_dir, exist_ok=True)
    enc, dec, ent = synth_state_dicts(valid_dim, seed)
    paths = [os.path.join(model_dir, "%s_%s.pt" % (prex, n)) for n in ("encoder", "decoder", "ent")]
    for p, d in zip(paths, (enc, dec, ent)):
        torch.save(d, p)
    merged = dict(enc)
    merged.update(dec)
    merged.update(ent)
    return paths, merged
