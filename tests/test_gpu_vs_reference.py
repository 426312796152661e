"""GPU tests against the UNMODIFIED reference extension (oracle/_ref/PCONV_ref.so, built by oracle/build_ref.py
from /root/reference/extension for sm_100).  Same inputs, same weights, both on the B200: every CUDA-core
operator must agree bit for bit, including the integer CDF tables and therefore the bitstream.
Skipped when the reference build is not present."""
import numpy as np
import pytest

from conftest import W64, smooth_images
from test_gpu_parity import N, T, assert_bit_equal

pytestmark = pytest.mark.gpu

WEIGHT = [float(v) for v in W64]


@pytest.fixture(scope="module")
def R(ref_ext, cuda):
    if ref_ext is None:
        pytest.skip("oracle/_ref/PCONV_ref.so not available")
    return ref_ext


@pytest.fixture(scope="module")
def P(cuda):
    from pseudocylindrical_convolution_b200 import PCONV
    return PCONV


def test_slice_uslice(R, P, cuda):
    for shape in [(1, 3, 512, 1024), (2, 4, 64, 128)]:
        x = T(smooth_images(*shape), cuda)
        mine = P.SphereSliceOp(16, 0, 0, WEIGHT, 0, False).forward(x)[0]
        ref = R.SphereSliceOp(16, 0, 0, WEIGHT, 0, False).forward(x)[0]
        assert_bit_equal(N(mine), N(ref), "slice %s" % (shape,))
        m2 = P.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False).forward(mine)[0]
        r2 = R.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False).forward(ref)[0]
        assert_bit_equal(N(m2), N(r2), "uslice %s" % (shape,))


@pytest.mark.parametrize("c,h,W,pad", [(3, 32, 1024, 1), (24, 16, 512, 2), (8, 4, 128, 1), (8, 2, 64, 2)])
def test_pad_fill(R, P, cuda, c, h, W, pad):
    import torch
    rng = np.random.default_rng(21)
    x = T(rng.standard_normal((16, c, h, W)).astype(np.float32), cuda)
    pm, pr = P.PseudoContextOp(16, 20, WEIGHT, 0, False), R.PseudoContextOp(16, 20, WEIGHT, 0, False)
    xm = P.PseudoFillOp(0, 16, 0, 0, pm.addr(), 0, 0, False).forward(x.clone())[0]
    xr = R.PseudoFillOp(0, 16, 0, 0, pr.addr(), 0, 0, False).forward(x.clone())[0]
    assert_bit_equal(N(xm), N(xr), "fill")
    ym = P.PseudoPadOp(pad, 16, pm.addr(), 0, False).forward(xm)[0]
    yr = R.PseudoPadOp(pad, 16, pr.addr(), 0, False).forward(xr)[0]
    assert_bit_equal(N(ym), N(yr), "pad")
    assert torch.isfinite(ym).all()


@pytest.mark.parametrize("version", [1, 0])
def test_entropy_pad(R, P, cuda, version):
    rng = np.random.default_rng(22)
    x = T(rng.standard_normal((16, 14, 4, 128)).astype(np.float32), cuda)
    cm = P.PseudoEntropyContextOp(16, 20, version, WEIGHT, 0, False)
    cr = R.PseudoEntropyContextOp(16, 20, version, WEIGHT, 0, False)
    xm = P.PseudoFillOp(0, 16, 0, 0, cm.addr(), 1, 0, False).forward(x.clone())[0]
    ym = P.PseudoEntropyPadOp(2, 16, cm.addr(), 0, False).forward(xm)[0]
    yr = R.PseudoEntropyPadOp(2, 16, cr.addr(), 0, False).forward(xm.clone())[0]
    assert_bit_equal(N(ym), N(yr), "entropy pad v%d" % version)


def test_quant_dquant_dtow(R, P, cuda):
    import torch
    rng = np.random.default_rng(23)
    Cc = 192
    theta = np.full((Cc, 8), np.log(1 / 9.0), np.float32) + rng.normal(scale=0.3, size=(Cc, 8)).astype(np.float32)
    theta[:, 0] = 1 / 9.0
    x = T(rng.random((16, Cc, 2, 64)).astype(np.float32), cuda)
    th = T(theta, cuda)
    cnt = torch.zeros((Cc, 8), device=cuda)
    cm, cr = P.PseudoContextOp(16, 20, WEIGHT, 0, False), R.PseudoContextOp(16, 20, WEIGHT, 0, False)
    vm, sm = P.PseudoQuantOp(Cc, 8, 16, 0.9, 100, 2, 0.1, cm.addr(), 0, False).forward(x, th, cnt, False)
    vr, sr = R.PseudoQuantOp(Cc, 8, 16, 0.9, 100, 2, 0.1, cr.addr(), 0, False).forward(x, th, cnt.clone(), False)
    assert_bit_equal(N(sm), N(sr), "symbols")
    assert_bit_equal(N(vm), N(vr), "dequantised")
    sub = sm[:, :56].contiguous()
    dm = P.DtowOp(2, True, 0, False).forward(sub)[0]
    dr = R.DtowOp(2, True, 0, False).forward(sub)[0]
    assert_bit_equal(N(dm), N(dr), "d2w")
    wm = P.DtowOp(2, False, 0, False).forward(dm)[0]
    wr = R.DtowOp(2, False, 0, False).forward(dr)[0]
    assert_bit_equal(N(wm), N(wr), "w2d")
    qm = P.PseudoDQuantOp(16, Cc, 8, cm.addr(), 0, False).forward(wm, th)[0]
    qr = R.PseudoDQuantOp(16, Cc, 8, cr.addr(), 0, False).forward(wr, th)[0]
    assert_bit_equal(N(qm), N(qr), "dquant")


def test_gmm_table(R, P, cuda):
    import torch
    rng = np.random.default_rng(24)
    n = 8192
    data = np.concatenate([rng.normal(size=(1, 3, 64, 128)) * 2, rng.normal(size=(1, 3, 64, 128)) * 1.5 + 0.5,
                           rng.random((1, 3, 64, 128)) * 9 - 4.5]).astype(np.float32)
    tn = torch.tensor([n], dtype=torch.int32)
    dm, dr = T(data, cuda), T(data, cuda)
    om = P.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward_batch(dm, tn)[0]
    orf = R.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward_batch(dr, tn)[0]
    assert_bit_equal(N(om), N(orf), "CDF tables")
    assert_bit_equal(N(dm), N(dr), "in-place softmax / delta")
    # plain (non-batch) entry point
    lg, dl, mu = (T(data[i].reshape(-1, 3), cuda).view(n // 128, 3, 1, 128).contiguous() for i in range(3))
    lg2, dl2, mu2 = lg.clone(), dl.clone(), mu.clone()
    a = P.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward(lg, dl, mu, tn)[0]
    b = R.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False).forward(lg2, dl2, mu2, tn)[0]
    assert_bit_equal(N(a)[:n], N(b)[:n], "CDF tables (plain)")


def _entropy_weights(G, seed=0):
    """Entropy-net weights drawn like the training net, stacked [logits, delta, mean]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    layers = []
    shapes = [(1, 3)] + [(3, 3)] * 10 + [(3, 3)]
    for li, (ci, co) in enumerate(shapes):
        fan_in = G * ci * 25
        w = torch.randn((3, G * co, G * ci, 5, 5), generator=g) * (2.0 / fan_in) ** 0.5 * 1.6
        b = torch.zeros((3, G * co))
        if li == len(shapes) - 1:
            b[1] = 2.0
        a = torch.full((3, G * co), 0.25)
        layers.append((w, b, a if li < len(shapes) - 1 else None))
    return layers


def _build_net(M, G, ctx_addr, layers, cuda):
    """The 12-layer DBT stack of pseudo_codec.py:79-87 out of raw ops of module M (product mirror or reference)."""
    ops = []
    for li, (w, b, a) in enumerate(layers):
        first, last = li == 0, li == len(layers) - 1
        gi = 1 if first else 3
        pad = M.EntropyCtxPadRun2Op(2, 16, G, first, ctx_addr, 0, False)
        conv = M.EntropyConv2Op(16, G * gi, G, G * 3, 5, 5 if first else 6, 2, 0 if last else 2, ctx_addr, 0, False)
        ops.append((pad, conv, w.to(cuda), b.to(cuda), a.to(cuda) if a is not None else None))
    adds = [M.EntropyAddOp(16, G * 3, G, 2, ctx_addr, 0, False) for _ in range(5)]
    return ops, adds


def _run_net(ops, adds, b):
    def layer(i, x):
        pad, conv, w, bias, act = ops[i]
        x = pad.forward(x)[0]
        return conv.forward_act_batch(x, w, bias, act)[0] if act is not None else conv.forward_batch(x, w, bias)[0]
    x = layer(0, b)
    for blk in range(5):
        y = layer(2 + 2 * blk, layer(1 + 2 * blk, x))
        x = adds[blk].forward(y, x)[0]
    return layer(11, x)


@pytest.mark.parametrize("G,h,W", [(14, 4, 128), (3, 2, 64)])
def test_full_wavefront_cdf_stream_and_bitstream(R, P, cuda, ref_coder, tmp_path, G, h, W):
    """The whole context model, stepped exactly like pseudo_codec.py:97-114 with both implementations: the CDF
    table of every step and the final bitstream must be identical; decoding with the product recovers the symbols."""
    import torch
    from pseudocylindrical_convolution_b200 import coder as mycoder
    layers = _entropy_weights(G)
    rng = np.random.default_rng(25)
    Hf = 16 * h
    sym = rng.integers(0, 8, size=(16, G, h, W)).astype(np.float32)
    sides = []
    for M in (P, R):
        ctx = M.EntropyContextOp(16, 18, WEIGHT, 0, False)
        a = ctx.addr()
        data = M.PseudoFillOp(0, 16, 0, 0, a, 2, 0, False).forward(T(sym, cuda))[0]
        ctx.start_context(W)
        ops, adds = _build_net(M, G, a, layers, cuda)
        sides.append(dict(ctx=ctx, data=data, ops=ops, adds=adds, ipt=M.DInput2Op(G, 16, 2, -3.5, 3, a, 0, False),
                          ext=M.DExtract2Op(16, G, True, a, 0, False), lab=M.DExtract2Op(16, G, True, a, 0, False),
                          gmm=M.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False),
                          label=torch.zeros((1, 1, Hf, W), device=cuda)))
    assert_bit_equal(N(sides[0]["data"]), N(sides[1]["data"]), "filled symbols")
    enc = mycoder.coder(str(tmp_path / "mine.bin"))
    enc.start_encoder()
    renc = None
    if ref_coder is not None:
        renc = ref_coder.coder(str(tmp_path / "ref.bin"))
        renc.start_encoder()
    tables, total = [], 0
    for s in range(Hf + W + G - 2):
        outs = []
        for sd in sides:
            b = sd["ipt"].forward(sd["label"])[0]
            y = _run_net(sd["ops"], sd["adds"], b)
            z, le = sd["ext"].forward_batch(y)
            vec = sd["gmm"].forward_batch(z, le)[0]
            ln = int(le[0].item())
            lab, _ = sd["lab"].forward(sd["data"])
            sd["label"] = lab
            outs.append((vec, ln, lab))
        (vm, lm, labm), (vr, lr, labr) = outs
        assert lm == lr
        if lm:
            assert_bit_equal(N(vm)[:lm], N(vr)[:lm], "CDF tables at step %d" % s)
            assert_bit_equal(N(labm).reshape(-1)[:lm], N(labr).reshape(-1)[:lm], "labels at step %d" % s)
            pred = vm[:lm].to(torch.int32).cpu()
            tl = labm.reshape(-1)[:lm].to(torch.int32).cpu()
            enc.encodes(pred, 8, tl, lm)
            if renc is not None:
                renc.encodes(vr.to(torch.int32).cpu(), 8, labr.to(torch.int32).cpu().view(-1), lr)
            tables.append(pred)
            total += lm
    enc.end_encoder()
    assert total == int((np.array([int(v / 64 * W + 0.5) for v in W64]) * h).sum()) * G
    mine = open(tmp_path / "mine.bin", "rb").read()
    if renc is not None:
        renc.end_encoder()
        assert mine == open(tmp_path / "ref.bin", "rb").read(), "bitstreams differ"
    assert len(mine) > 16
