"""Golden cases shared by the generator (tools/make_golden.py, run ON the B200 box with the UNMODIFIED reference
extension oracle/_ref/PCONV_ref.so), the CPU tests (oracle vs golden) and the GPU tests (product vs golden).

Every case is a function of a PCONV-shaped module `M` (the reference extension or the product mirror - they expose
the same classes, extension/main.cpp:4-137) returning an ordered dict of numpy arrays.  Inputs are drawn from
numpy's PCG64 with fixed seeds and quantised to values that are exact in float32, so the .npz files only
need to hold the OUTPUTS (plus the bitstream)."""
from collections import OrderedDict

import numpy as np

W64 = [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]
WEIGHT = [float(v) for v in W64]


def _grid(rng, shape, lo=-2.0, hi=2.0, den=256):
    """Random values on a 1/den grid (exact in fp32, stable across numpy versions as integers)."""
    return (rng.integers(int(lo * den), int(hi * den) + 1, size=shape).astype(np.float32) / np.float32(den)).astype(np.float32)


# ------------------------------------------------------------------------------------------------ inputs
def inputs_tiles():
    rng = np.random.default_rng(101)
    return dict(erp=_grid(rng, (1, 2, 32, 64), 0, 1), tiles=_grid(rng, (16, 2, 2, 64)), ctiles=_grid(rng, (16, 3, 2, 64), -3.5, 3.5, 2),
                d2w=_grid(rng, (2, 8, 3, 5)))


def inputs_quant():
    rng = np.random.default_rng(102)
    theta = (np.float32(np.log(1 / 9.0)) + _grid(rng, (8, 8), -0.5, 0.5, 64)).astype(np.float32)
    theta[:, 0] = np.float32(1 / 9.0)
    return dict(x=_grid(rng, (16, 8, 2, 64), 0, 1, 4096), theta=theta)


def inputs_gmm():
    rng = np.random.default_rng(103)
    n = 512
    logit = _grid(rng, (n, 3), -4, 4)
    delta = _grid(rng, (n, 3), -1, 4)
    mean = _grid(rng, (n, 3), -4.5, 4.5)
    label = (rng.integers(0, 8, size=(n, 1)).astype(np.float32) - np.float32(3.5))
    return dict(logit=logit, delta=delta, mean=mean, label=label, n=n)


def inputs_wavefront(G=3, h=2, W=64):
    rng = np.random.default_rng(104)
    layers = []
    shapes = [(1, 3)] + [(3, 3)] * 11
    for li, (ci, co) in enumerate(shapes):
        fan_in = G * ci * 25
        scale = np.float32(np.sqrt(2.0 / fan_in) * 1.6 / 64.0)
        w = rng.integers(-127, 128, size=(3, G * co, G * ci, 5, 5)).astype(np.float32) * scale
        b = np.zeros((3, G * co), np.float32)
        if li == len(shapes) - 1:
            b[1] = 2.0
        a = np.full((3, G * co), 0.25, np.float32) if li < len(shapes) - 1 else None
        layers.append((w.astype(np.float32), b, a))
    sym = rng.integers(0, 8, size=(16, G, h, W)).astype(np.float32)
    return dict(G=G, h=h, W=W, layers=layers, sym=sym)


# ------------------------------------------------------------------------------------------------ module-driven cases
def _T(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _N(t):
    return t.detach().cpu().numpy().copy()


def case_tiles(M, dev):
    i = inputs_tiles()
    out = OrderedDict()
    sl = M.SphereSliceOp(16, 0, 0, WEIGHT, 0, False).forward(_T(i["erp"], dev))[0]
    out["slice"] = _N(sl)
    out["uslice"] = _N(M.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False).forward(sl)[0])
    ctx = M.PseudoContextOp(16, 20, WEIGHT, 0, False)
    filled = M.PseudoFillOp(0, 16, 0, 0, ctx.addr(), 0, 0, False).forward(_T(i["tiles"], dev))[0]
    out["fill"] = _N(filled)
    for p in (1, 2):
        out["pad%d" % p] = _N(M.PseudoPadOp(p, 16, ctx.addr(), 0, False).forward(filled)[0])
    padded = _T(out["pad2"], dev)
    out["fill_pad2_trim1"] = _N(M.PseudoFillOp(2, 16, 0, 1, ctx.addr(), 0, 0, False).forward(padded.clone())[0])
    for v in (1, 0):
        ectx = M.PseudoEntropyContextOp(16, 20, v, WEIGHT, 0, False)
        cf = M.PseudoFillOp(0, 16, 0, 0, ectx.addr(), 1, 0, False).forward(_T(i["ctiles"], dev))[0]
        out["entropy_pad_v%d" % v] = _N(M.PseudoEntropyPadOp(2, 16, ectx.addr(), 0, False).forward(cf)[0])
    d = M.DtowOp(2, True, 0, False).forward(_T(i["d2w"], dev))[0]
    out["d2w"] = _N(d)
    out["w2d"] = _N(M.DtowOp(2, False, 0, False).forward(d)[0])
    return out


def case_quant(M, dev):
    import torch
    i = inputs_quant()
    out = OrderedDict()
    ctx = M.PseudoContextOp(16, 20, WEIGHT, 0, False)
    th = _T(i["theta"], dev)
    cnt = torch.zeros((8, 8), device=dev)
    val, sym = M.PseudoQuantOp(8, 8, 16, 0.9, 100, 2, 0.1, ctx.addr(), 0, False).forward(_T(i["x"], dev), th, cnt, False)
    out["val"], out["sym"] = _N(val), _N(sym)
    out["dquant"] = _N(M.PseudoDQuantOp(16, 8, 8, ctx.addr(), 0, False).forward(sym.contiguous(), th)[0])
    return out


def case_gmm(M, dev):
    import torch
    i = inputs_gmm()
    n = i["n"]
    out = OrderedDict()
    data = np.stack([i["logit"].reshape(-1), i["delta"].reshape(-1), i["mean"].reshape(-1)]).reshape(3, 3, 8, 64)
    d = _T(data, dev)
    tn = torch.tensor([n], dtype=torch.int32)
    out["cdf_batch"] = _N(M.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward_batch(d, tn)[0]).astype(np.int32)
    out["softmax"] = _N(d).reshape(3, -1)[0].reshape(n, 3)
    out["delta_clamped"] = _N(d).reshape(3, -1)[1].reshape(n, 3)
    lg, dl, mu = (_T(data[k].reshape(n // 64, 3, 1, 64), dev) for k in range(3))
    out["cdf_plain"] = _N(M.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward(lg, dl, mu, tn)[0])[:n].astype(np.int32)
    w = _T(out["softmax"], dev)
    out["nll"] = _N(M.EntropyGmmOp(3, 0, 0, False).forward(w, _T(out["delta_clamped"], dev), _T(i["mean"], dev), _T(i["label"], dev))[0])
    return out


def _fold(a):
    """order-sensitive 64-bit digest of a float32 array's bit patterns"""
    import hashlib
    raw = np.ascontiguousarray(a, np.float32).tobytes()
    return np.uint64(int.from_bytes(hashlib.blake2b(raw, digest_size=8).digest(), "little"))


def case_wavefront(M, dev, coder_mod=None, tmp_path=None):
    """The whole context model stepped like pseudo_codec.py:97-114: per-step symbol counts, a digest of the GMM
    parameters extracted at each step, every CDF table row, and (when a coder module is given) the bitstream."""
    import torch
    i = inputs_wavefront()
    G, h, W = i["G"], i["h"], i["W"]
    Hf = 16 * h
    ctx = M.EntropyContextOp(16, 18, WEIGHT, 0, False)
    a = ctx.addr()
    data = M.PseudoFillOp(0, 16, 0, 0, a, 2, 0, False).forward(_T(i["sym"], dev))[0]
    ctx.start_context(W)
    ops = []
    for li, (w, b, act) in enumerate(i["layers"]):
        first, last = li == 0, li == len(i["layers"]) - 1
        gi = 1 if first else 3
        pad = M.EntropyCtxPadRun2Op(2, 16, G, first, a, 0, False)
        conv = M.EntropyConv2Op(16, G * gi, G, G * 3, 5, 5 if first else 6, 2, 0 if last else 2, a, 0, False)
        ops.append((pad, conv, _T(w, dev), _T(b, dev), _T(act, dev) if act is not None else None))
    adds = [M.EntropyAddOp(16, G * 3, G, 2, a, 0, False) for _ in range(5)]
    ipt = M.DInput2Op(G, 16, 2, -3.5, 3, a, 0, False)
    ext = M.DExtract2Op(16, G, True, a, 0, False)
    lab = M.DExtract2Op(16, G, True, a, 0, False)
    gmm = M.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False)

    def layer(k, x):
        pad, conv, w, bias, act = ops[k]
        x = pad.forward(x)[0]
        return conv.forward_act_batch(x, w, bias, act)[0] if act is not None else conv.forward_batch(x, w, bias)[0]

    enc = None
    if coder_mod is not None:
        enc = coder_mod.coder(str(tmp_path))
        enc.start_encoder()
    label = torch.zeros((1, 1, Hf, W), device=dev)
    counts, digests, tables, labels = [], [], [], []
    for s in range(Hf + W + G - 2):
        x = layer(0, ipt.forward(label)[0])
        for blk in range(5):
            y = layer(2 + 2 * blk, layer(1 + 2 * blk, x))
            x = adds[blk].forward(y, x)[0]
        z, le = ext.forward_batch(layer(11, x))
        ln = int(le[0].item())
        zz = _N(z).reshape(3, -1)[:, :ln * 3]
        vec = gmm.forward_batch(z, le)[0]
        label, _ = lab.forward(data)
        counts.append(ln)
        digests.append(_fold(zz) if ln else np.uint64(0))
        if ln:
            pred = vec[:ln].to(torch.int32).cpu()
            tl = label.reshape(-1)[:ln].to(torch.int32).cpu()
            tables.append(pred.numpy().copy())
            labels.append(tl.numpy().copy())
            if enc is not None:
                enc.encodes(pred.contiguous(), 8, tl.contiguous(), ln)
    out = OrderedDict(counts=np.array(counts, np.int32), digests=np.array(digests, np.uint64),
                      tables=np.concatenate(tables).astype(np.int32), labels=np.concatenate(labels).astype(np.uint8))
    if enc is not None:
        enc.end_encoder()
        out["bitstream"] = np.frombuffer(open(str(tmp_path), "rb").read(), np.uint8).copy()
    return out


CASES = OrderedDict(tiles=case_tiles, quant=case_quant, gmm=case_gmm, wavefront=case_wavefront)
