"""GPU parity tests: every CUDA entry point (called through the PCONV mirror, i.e. through the C ABI) against
the CPU oracle on the same seeded inputs - bit-exact for gathers, integer work and fixed-order float
reductions; stated tolerances only where libm (expf/erff) or an unpinned accumulation order is involved."""
import numpy as np
import pytest

from conftest import W64, smooth_images

pytestmark = pytest.mark.gpu

WEIGHT = [float(v) for v in W64]


def T(a, dev):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bit_equal(got, want, what=""):
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    if got.dtype.kind == "f":
        same = bits(got) == bits(want)
    else:
        same = got == want
    if not same.all():
        bad = np.argwhere(~same)
        i = tuple(bad[0])
        raise AssertionError("%s: %d / %d elements differ, first at %s: got %r want %r" %
                             (what, len(bad), same.size, i, got[i], want[i]))


@pytest.fixture(scope="module")
def P(cuda):
    from pseudocylindrical_convolution_b200 import PCONV
    return PCONV


# ------------------------------------------------------------------------------------------------ tables
@pytest.mark.parametrize("W", [64, 128, 1024, 2048])
def test_cubic_tables(P, cuda, orc, W):
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call, int_array
    wl = orc.band_widths(W64, 64, W)
    for fn, ofn in (("pcx_slice_table", orc.slice_table), ("pcx_uslice_table", orc.uslice_table)):
        src = torch.zeros((16, W), dtype=torch.int32, device=cuda)
        wt = torch.zeros((16, W, 4), dtype=torch.float32, device=cuda)
        call(fn, int_array(wl), 16, W, C.c_void_p(src.data_ptr()), C.c_void_p(wt.data_ptr()), None)
        torch.cuda.synchronize()
        osrc, owt = ofn(wl, W)
        assert_bit_equal(N(src), osrc, fn + " src")
        assert_bit_equal(N(wt), owt, fn + " weights")


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("h,W,pad", [(4, 128, 2), (2, 64, 1), (16, 512, 1), (1, 64, 2), (32, 1024, 2)])
def test_halo_tables(P, cuda, orc, mode, h, W, pad):
    ctx = (P.PseudoContextOp(16, 20, WEIGHT, 0, False) if mode == 0 else
           P.PseudoEntropyContextOp(16, 20, 1 if mode == 1 else 0, WEIGHT, 0, False))
    band, row, col, tw = ctx.halo(3, h, W, pad)
    ob, orow, ocol, otw = orc.halo_table(orc.band_widths(W64, 16 * h, W), h, W, pad, mode)
    assert_bit_equal(N(band), ob, "band")
    assert_bit_equal(N(row), orow, "row")
    assert_bit_equal(N(col), ocol, "col")
    assert_bit_equal(N(tw), otw, "weight")


# ------------------------------------------------------------------------------------------------ tile pipeline
@pytest.mark.parametrize("n,c,H,W", [(1, 3, 64, 128), (2, 3, 512, 1024), (1, 5, 32, 64), (1, 2, 256, 2048), (1, 1, 128, 4096)])
def test_slice_uslice_bit_exact(P, cuda, orc, n, c, H, W):
    x = smooth_images(n, c, H, W)
    wl = orc.band_widths(W64, H, W)
    sl = P.SphereSliceOp(16, 0, 0, WEIGHT, 0, False)
    tiles = sl.forward(T(x, cuda))[0]
    want = orc.sphere_slice(x, wl)
    assert_bit_equal(N(tiles), want, "slice")
    us = P.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False)
    erp = us.forward(tiles)[0]
    assert_bit_equal(N(erp), orc.sphere_uslice(want, wl), "uslice")
    # property: bands are a smooth resampling, the round trip stays close to the input
    assert np.abs(N(erp) - x).mean() < 0.05


def test_slice_with_pad_writes_interior_only(P, cuda, orc):
    import torch
    x = smooth_images(1, 2, 64, 128)
    wl = orc.band_widths(W64, 64, 128)
    sl = P.SphereSliceOp(16, 0, 1, WEIGHT, 0, False)
    out = sl.forward(T(x, cuda))[0]
    assert tuple(out.shape) == (16, 2, 6, 130)
    want = orc.sphere_slice(x, wl, pad=1)
    assert_bit_equal(N(out)[:, :, 1:-1, 1:-1], want[:, :, 1:-1, 1:-1], "slice pad=1 interior")
    us = P.SphereUsliceOp(16, 0, 1, WEIGHT, 0, False)
    assert_bit_equal(N(us.forward(out)[0]), orc.sphere_uslice(want, wl, pad=1), "uslice pad=1")


@pytest.mark.parametrize("c,h,W,pad", [(3, 4, 128, 1), (3, 4, 128, 2), (8, 16, 512, 1), (4, 2, 64, 2), (2, 32, 1024, 2), (2, 64, 2048, 1)])
def test_pad_bit_exact(P, cuda, orc, c, h, W, pad):
    rng = np.random.default_rng(7)
    wl = orc.band_widths(W64, 16 * h, W)
    x = orc.pseudo_fill(rng.standard_normal((32, c, h, W)).astype(np.float32), wl)     # 2 images
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    op = P.PseudoPadOp(pad, 16, ctx.addr(), 0, False)
    got = N(op.forward(T(x, cuda))[0])
    assert_bit_equal(got, orc.pseudo_pad(x, wl, pad), "pad")


@pytest.mark.parametrize("version", [1, 0])
@pytest.mark.parametrize("c,h,W,pad", [(14, 4, 128, 2), (3, 2, 64, 2), (6, 8, 256, 1)])
def test_entropy_pad_bit_exact(P, cuda, orc, version, c, h, W, pad):
    rng = np.random.default_rng(8)
    wl = orc.band_widths(W64, 16 * h, W)
    x = orc.pseudo_fill(rng.standard_normal((16, c, h, W)).astype(np.float32), wl)
    ctx = P.PseudoEntropyContextOp(16, 20, version, WEIGHT, 0, False)
    op = P.PseudoEntropyPadOp(pad, 16, ctx.addr(), 0, False)
    got = N(op.forward(T(x, cuda))[0])
    assert_bit_equal(got, orc.pseudo_entropy_pad(x, wl, pad, version), "entropy pad v%d" % version)


def test_halo_fill_matches_pad(P, cuda, orc):
    """The in-place halo refresh used by the fused transforms produces the same padded tile as the copy."""
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call, int_array
    rng = np.random.default_rng(9)
    c, h, W, pad, pitch = 5, 8, 256, 1, 260
    wl = orc.band_widths(W64, 16 * h, W)
    x = orc.pseudo_fill(rng.standard_normal((16, c, h, W)).astype(np.float32), wl)
    want = orc.pseudo_pad(x, wl, pad)
    buf = np.zeros((16, c, h + 2 * pad, pitch), np.float32)
    buf[:, :, pad:pad + h, pad:pad + W] = x
    d = T(buf, cuda)
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    band, row, col, tw = ctx.halo(c, h, W, pad)
    call("pcx_halo_fill", C.c_void_p(d.data_ptr()), 1, c, h, W, 16, pad, int_array(wl), C.c_void_p(band.data_ptr()),
         C.c_void_p(row.data_ptr()), C.c_void_p(col.data_ptr()), C.c_void_p(tw.data_ptr()), pitch, None)
    torch.cuda.synchronize()
    assert_bit_equal(N(d)[:, :, :, :W + 2 * pad], want, "halo_fill")
    # pitched full pad
    out = torch.full((16, c, h + 2 * pad, pitch), 7.0, device=cuda)
    call("pcx_pad_fwd", C.c_void_p(T(x, cuda).data_ptr()), C.c_void_p(out.data_ptr()), 1, c, h, W, 16, pad, int_array(wl),
         C.c_void_p(band.data_ptr()), C.c_void_p(row.data_ptr()), C.c_void_p(col.data_ptr()), C.c_void_p(tw.data_ptr()), pitch, None)
    torch.cuda.synchronize()
    assert_bit_equal(N(out)[:, :, :, :W + 2 * pad], want, "pitched pad")
    assert (N(out)[:, :, :, W + 2 * pad:] == 0).all()


@pytest.mark.parametrize("pad,trim,fvalue", [(0, 0, 0), (2, 0, 0), (2, 1, 0), (0, 0, 1)])
def test_fill_bit_exact(P, cuda, orc, pad, trim, fvalue):
    rng = np.random.default_rng(10)
    h, W = 4, 128
    x = rng.standard_normal((32, 3, h + 2 * pad, W + 2 * pad)).astype(np.float32)
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    op = P.PseudoFillOp(pad, 16, fvalue, trim, ctx.addr(), 0, 0, False)
    d = T(x, cuda)
    out = op.forward(d)[0]
    assert out.data_ptr() == d.data_ptr()                      # in place
    wl = orc.band_widths(W64, 16 * (h + 2 * pad), W + 2 * pad)  # the reference derives widths from the tensor's own extent
    assert_bit_equal(N(out), orc.pseudo_fill(x, wl, pad, trim, float(fvalue)), "fill")


def test_dtow_bit_exact(P, cuda, orc):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((16, 56, 2, 64)).astype(np.float32)
    up = P.DtowOp(2, True, 0, False).forward(T(x, cuda))[0]
    assert_bit_equal(N(up), orc.dtow(x, 2, True), "d2w")
    down = P.DtowOp(2, False, 0, False).forward(up)[0]
    assert_bit_equal(N(down), x, "w2d(d2w(x))")


# ------------------------------------------------------------------------------------------------ quantiser
def test_quant_dquant(P, cuda, orc):
    import torch
    rng = np.random.default_rng(12)
    Cc, h, W = 192, 2, 64
    wl = orc.band_widths(W64, 16 * h, W)
    theta = np.full((Cc, 8), np.log(1 / 9.0), np.float32) + rng.normal(scale=0.2, size=(Cc, 8)).astype(np.float32)
    theta[:, 0] = 1 / 9.0 + rng.normal(scale=0.02, size=Cc).astype(np.float32)
    x = rng.random((16, Cc, h, W)).astype(np.float32) * 1.2 - 0.1
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    q = P.PseudoQuantOp(Cc, 8, 16, 0.9, 100, 2, 0.1, ctx.addr(), 0, False)
    val, sym = q.forward(T(x, cuda), T(theta, cuda), torch.zeros((Cc, 8), device=cuda), False)
    steps_gpu = N(q._outs["steps"])
    # step table: expf of libdevice vs glibc, <= 2 ulp
    np.testing.assert_allclose(steps_gpu, orc.quant_steps(theta), rtol=3e-7, atol=0)
    oval, osym, _ = orc.pseudo_quant(x, steps_gpu, wl)           # same table -> the search must agree exactly
    assert_bit_equal(N(sym), osym, "symbols")
    assert_bit_equal(N(val), oval, "dequantised values")
    dq = P.PseudoDQuantOp(16, Cc, 8, ctx.addr(), 0, False)
    sub = np.ascontiguousarray(osym[:, :56])
    rec = dq.forward(T(sub, cuda), T(theta, cuda))[0]
    cen_gpu = N(dq._outs["centres"])
    np.testing.assert_allclose(cen_gpu, orc.dquant_centres(theta)[:56], rtol=1e-6, atol=0)
    assert_bit_equal(N(rec), orc.pseudo_dquant(sub, cen_gpu, wl), "dquant lookups")
    # quantise -> dequantise lands on the nearest centre
    np.testing.assert_allclose(N(rec), oval[:, :56], rtol=0, atol=2e-6)


# ------------------------------------------------------------------------------------------------ dense
def _conv_desc(NN, Ci, Hi, Wi, Co, k, s, act, wl_out, impl):
    from pseudocylindrical_convolution_b200._lib import ConvDesc
    d = ConvDesc()
    Ho, Wo = (Hi - k) // s + 1, (Wi - k) // s + 1
    d.N, d.npart, d.Ci, d.Hi, d.in_pitch = NN // 16, 16, Ci, Hi, Wi
    d.Co, d.Ho, d.Wo, d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = Co, Ho, Wo, Ho, Wo, 0, 0
    d.k, d.stride, d.act, d.impl = k, s, act, impl
    d.aux_rows, d.aux_pitch, d.aux_y0, d.aux_x0 = Ho, Wo, 0, 0
    for g in range(16):
        d.wl_out[g] = min(int(wl_out[g]), Wo)
    return d, Ho, Wo


@pytest.mark.parametrize("Ci,Co,k,s,act", [(3, 16, 3, 2, 1), (24, 40, 3, 1, 1), (32, 32, 1, 1, 0), (16, 24, 1, 2, 0), (24, 24, 1, 1, 2)])
def test_conv_direct_vs_oracle(cuda, orc, Ci, Co, k, s, act):
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call
    rng = np.random.default_rng(13)
    h, W = 8, 128
    Hi, Wi = h + (2 if k == 3 else 0), W + (2 if k == 3 else 0)
    wl = orc.band_widths(W64, 16 * (h // s), W // s)
    x = rng.standard_normal((16, Ci, Hi, Wi)).astype(np.float32)
    w = (rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    slope = rng.random(Co).astype(np.float32) * 0.5
    d, Ho, Wo = _conv_desc(16, Ci, Hi, Wi, Co, k, s, act, wl, 1)
    res = rng.standard_normal((16, Co, Ho, Wo)).astype(np.float32)
    mul = rng.standard_normal((16, Co, Ho, Wo)).astype(np.float32)
    y = torch.empty((16, Co, Ho, Wo), device=cuda)
    dx, dw, db, ds, dm, dr = (T(a, cuda) for a in (x, w, b, slope, mul, res))
    call("pcx_conv2d_fwd", C.byref(d), *(C.c_void_p(t.data_ptr()) for t in (dx, dw, db, ds, dm, dr, y)), None)
    torch.cuda.synchronize()
    ref = orc.conv2d(x, w, b, s)
    if act == 1:
        ref = np.where(ref < 0, ref * slope[None, :, None, None], ref)
    elif act == 2:
        ref = 1 / (1 + np.exp(-ref.astype(np.float64)))
    ref = (res + mul * ref).astype(np.float32)
    for g in range(16):
        ref[g, :, :, d.wl_out[g]:] = 0
    np.testing.assert_allclose(N(y), ref, rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_vs_oracle(cuda, orc, inverse):
    import torch
    from pseudocylindrical_convolution_b200.PCONV_operator import PseudoContextV2, PseudoGDNV2
    rng = np.random.default_rng(14)
    h, W = 4, 128
    wl = orc.band_widths(W64, 16 * h, W)
    ctx = PseudoContextV2(16, True, device=0)
    gdn = PseudoGDNV2(192, 16, ctx, 0, inverse=inverse)
    with torch.no_grad():
        gdn.gamma.add_(torch.rand(192, 192, device=cuda) * 0.02)
        gdn.beta.add_(torch.rand(192, device=cuda) * 0.1)
    x = rng.standard_normal((16, 192, h, W)).astype(np.float32)
    res = rng.standard_normal((16, 192, h, W)).astype(np.float32)
    want = orc.gdn(x, N(gdn.beta), N(gdn.gamma), wl, inverse)
    np.testing.assert_allclose(N(gdn(T(x, cuda))), want, rtol=3e-5, atol=1e-6)
    want_res = want + orc.pseudo_fill(res, wl)
    np.testing.assert_allclose(N(gdn(T(x, cuda), residual=T(res, cuda))), want_res, rtol=3e-5, atol=2e-6)


# ------------------------------------------------------------------------------------------------ context model
def _ctx_ops(P, G, gi, go, pad_in, pad_out, constrain, input_flag):
    ctx = P.EntropyContextOp(16, 18, WEIGHT, 0, False)
    a = ctx.addr()
    return (ctx, P.EntropyCtxPadRun2Op(2, 16, G, input_flag, a, 0, False),
            P.EntropyConv2Op(16, G * gi, G, G * go, 5, constrain, pad_in, pad_out, a, 0, False),
            P.EntropyAddOp(16, G * go, G, 2, a, 0, False))


@pytest.mark.parametrize("G,h,W,nimg", [(3, 2, 64, 1), (14, 4, 128, 1), (4, 2, 64, 2)])
def test_wavefront_ops_bit_exact(P, cuda, orc, G, h, W, nimg):
    """ctx pad, masked conv (+PReLU), add, d_input, d_extract stepped side by side with the oracle for a whole image."""
    import torch
    rng = np.random.default_rng(15)
    wl = orc.band_widths(W64, 16 * h, W)
    geom = orc.CtxGeom(wl, h, W, 2)
    nb, gi, go = 3, 1, 3
    ctx, pad_in, conv1, _ = _ctx_ops(P, G, gi, go, 2, 2, 5, True)
    _, pad_h, conv2, add = _ctx_ops(P, G, go, go, 2, 2, 6, False)
    # one shared context so both chains see the same tables
    a = ctx.addr()
    pad_h = P.EntropyCtxPadRun2Op(2, 16, G, False, a, 0, False)
    conv2 = P.EntropyConv2Op(16, G * go, G, G * go, 5, 6, 2, 0, a, 0, False)
    add = P.EntropyAddOp(16, G * go, G, 2, a, 0, False)
    ipt = P.DInput2Op(G, 16, 2, -3.5, 3, a, 0, False)
    ext = P.DExtract2Op(16, G, True, a, 0, False)
    lab = P.DExtract2Op(16, G, True, a, 0, False)

    w1 = (rng.standard_normal((nb, G * go, G * gi, 5, 5)) * 0.3).astype(np.float32)
    b1 = rng.standard_normal((nb, G * go)).astype(np.float32) * 0.1
    a1 = rng.random((nb, G * go)).astype(np.float32)
    w2 = (rng.standard_normal((nb, G * go, G * go, 5, 5)) * 0.1).astype(np.float32)
    b2 = rng.standard_normal((nb, G * go)).astype(np.float32) * 0.1
    dw1, db1, da1, dw2, db2 = (T(v, cuda) for v in (w1, b1, a1, w2, b2))
    data = orc.pseudo_fill(rng.integers(0, 8, size=(nimg * 16, G, h, W)).astype(np.float32), wl)
    ddata = T(data, cuda)

    Hf = 16 * h
    NN = nb * nimg * 16
    o_in = np.zeros((NN, G, h + 4, W + 4), np.float32)
    o_mid = np.zeros((NN, G * go, h + 4, W + 4), np.float32)
    o_out = np.zeros((NN, G * go, h, W), np.float32)
    o_ext = np.zeros((nb * nimg, go, Hf, W), np.float32)
    o_lab = np.zeros((nimg, 1, Hf, W), np.float32)
    prev = np.zeros((nimg, 1, Hf, W), np.float32)
    nsteps = geom.nsteps(G)
    checked = 0
    for s in range(nsteps):
        # ---- product
        b = ipt.forward(T(prev, cuda))[0]
        b = pad_in.forward(b)[0]
        m = conv1.forward_act_batch(b, dw1, db1, da1)[0]
        m_res = m.clone()
        m = pad_h.forward(m)[0]
        y = conv2.forward_batch(m, dw2, db2)[0]
        m_res = add.forward(m_res, m)[0]
        z, cnt = ext.forward_batch(y)
        l, lcnt = lab.forward(ddata)
        # ---- oracle
        orc.dinput_step(prev.reshape(-1), o_in, geom, G, nimg, 2, -3.5, 3, s)
        orc.ctx_pad_step(o_in, geom, G, s - 1)
        orc.ctx_conv_step(o_in, w1, b1, a1, o_mid, geom, G, nimg, 2, 2, 5, s)
        o_res = o_mid.copy()
        orc.ctx_pad_step(o_mid, geom, G, s)
        orc.ctx_conv_step(o_mid, w2, b2, None, o_out, geom, G, nimg, 2, 0, 6, s)
        orc.ctx_add_step(o_res, o_mid, geom, G, 2, s)
        n_ext = orc.dextract_step(o_out, o_ext, geom, G, s, True)
        n_lab = orc.dextract_step(data, o_lab, geom, G, s, False)
        if s % 7 == 0 or s > nsteps - 4 or s < 3:
            assert_bit_equal(N(b), o_in, "step %d input + ctx pad" % s)
            assert_bit_equal(N(m), o_mid, "step %d conv1 + pad" % s)
            assert_bit_equal(N(y), o_out, "step %d conv2" % s)
            assert_bit_equal(N(m_res), o_res, "step %d add" % s)
            checked += 1
        assert int(cnt[0]) == n_ext and int(lcnt[0]) == n_lab
        if n_ext:
            zz = N(z).reshape(nb, -1)[:, :n_ext * go]
            assert_bit_equal(zz, o_ext.reshape(nb, -1)[:, :n_ext * go], "step %d extract" % s)
            assert_bit_equal(N(l).reshape(-1)[:n_lab], o_lab.reshape(-1)[:n_lab], "step %d labels" % s)
        prev = np.zeros((nimg, 1, Hf, W), np.float32)
        prev.reshape(-1)[:n_lab] = o_lab.reshape(-1)[:n_lab]
    assert checked > 5
    # after the last step every valid cell of every group has been produced exactly once
    for g in range(16):
        assert (o_in[g, :, 2:-2, 2:2 + wl[g]] != 0).any()


def test_gmm_table_vs_oracle(P, cuda, orc):
    import torch
    rng = np.random.default_rng(16)
    n = 8192
    logit = rng.normal(size=(n, 3)).astype(np.float32) * 2
    delta = (rng.normal(size=(n, 3)) * 1.5 + 0.5).astype(np.float32)
    mean = (rng.random((n, 3)) * 9 - 4.5).astype(np.float32)
    data = np.stack([logit.reshape(-1), delta.reshape(-1), mean.reshape(-1)]).reshape(3, 3, 64, 128)
    op = P.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False)
    d = T(data, cuda)
    out = N(op.forward_batch(d, torch.tensor([n], dtype=torch.int32))[0])
    cdf, w, dl = orc.gmm_table(logit, delta, mean)
    assert out.shape == (n, 9)
    assert (out == np.round(out)).all()
    got = out.astype(np.int64)
    assert (got[:, 0] == 0).all() and (got[:, 8] == 65536).all() and (np.diff(got, axis=1) > 0).all()
    diff = np.abs(got - cdf)
    assert diff.max() <= 1, "CDF entries may differ from the CPU oracle by at most one count (libm erff/expf)"
    assert (diff > 0).mean() < 0.02
    # in-place side effects: softmax weights and clamped deltas (entropy_gmm_table_cuda.cu:29-56)
    np.testing.assert_allclose(N(d).reshape(3, -1)[0].reshape(n, 3), w, rtol=2e-6, atol=1e-8)
    assert_bit_equal(N(d).reshape(3, -1)[1].reshape(n, 3), dl, "delta clamp")
    # partial count: rows beyond tn are left untouched
    d2 = T(data, cuda)
    op2 = P.EntropyGmmTableOp(8, 3.5, 3, 65536.0, 1e-6, 0, False)
    out2 = op2.forward_batch(d2, torch.tensor([100], dtype=torch.int32))[0]
    assert_bit_equal(N(out2)[:100], out[:100], "first rows")
    assert_bit_equal(N(d2).reshape(3, -1)[0][300:], data.reshape(3, -1)[0][300:], "untouched logits")


def test_gmm_nll_vs_oracle(P, cuda, orc):
    rng = np.random.default_rng(17)
    n = 5000
    w = rng.random((n, 3)).astype(np.float32)
    w /= w.sum(1, keepdims=True)
    delta = (rng.random((n, 3)) * 2 + 0.05).astype(np.float32)
    mean = (rng.random((n, 3)) * 8 - 4).astype(np.float32)
    label = (rng.integers(0, 8, size=(n, 1)) - 3.5).astype(np.float32)
    op = P.EntropyGmmOp(3, 0, 0, False)
    got = N(op.forward(T(w, cuda), T(delta, cuda), T(mean, cuda), T(label, cuda))[0])
    want = orc.gmm_nll(w, delta, mean, label.reshape(-1))
    # the loss is -log(p + 1e-7) with p a difference of two float erf values: compare on the probability scale
    # (float32 cancellation noise ~1e-7 dominates the tails), and tightly where p is not tiny
    np.testing.assert_allclose(np.exp(-got.astype(np.float64)), np.exp(-want.astype(np.float64)), rtol=1e-5, atol=3e-7)
    big = want < 8
    np.testing.assert_allclose(got[big], want[big], rtol=2e-3, atol=2e-5)


def test_errors_are_loud(P, cuda):
    import torch
    from pseudocylindrical_convolution_b200._lib import PcxError
    sl = P.SphereSliceOp(16, 0, 0, WEIGHT, 0, False)
    with pytest.raises(PcxError):
        sl.forward(torch.zeros((1, 3, 65, 128), device=cuda))           # height not a multiple of npart
    with pytest.raises(TypeError):
        sl.forward(torch.zeros((1, 3, 64, 128), device=cuda, dtype=torch.float64))
    with pytest.raises(TypeError):
        sl.forward(torch.zeros((1, 3, 64, 128)))                        # CPU tensor: no CPU fallback
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    with pytest.raises(PcxError):
        P.PseudoPadOp(12, 16, ctx.addr(), 0, False).forward(torch.zeros((16, 3, 16, 128), device=cuda))
    with pytest.raises(PcxError):
        P.PseudoPadOp(1, 16, "0xdeadbeef", 0, False)
    with pytest.raises(NotImplementedError):
        sl.backward(torch.zeros((16, 3, 4, 128), device=cuda))


# ------------------------------------------------------------------------------------------------ tensor-core conv
@pytest.mark.parametrize("Ci,Co,k,s,act,h,W", [
    (96, 96, 3, 1, 1, 8, 128),      # ResidualBlock conv2
    (192, 192, 3, 1, 1, 8, 128),    # ResidualBlockV2
    (192, 96, 1, 1, 1, 10, 130),    # ResidualBlock conv1 on the padded tile (odd pitch)
    (96, 192, 1, 1, 0, 8, 128),     # ResidualBlock conv3 (+ residual)
    (192, 768, 3, 1, 1, 4, 64),     # ResidualBlockUp conv1 (4 N-tiles)
    (192, 12, 3, 1, 0, 8, 128),     # last decoder layer (N padded to 16)
    (192, 192, 3, 2, 1, 8, 128),    # ResidualBlockDown conv1 (TMA element stride 2)
    (192, 192, 1, 2, 0, 8, 128),    # ResidualBlockDown shortcut
    (192, 192, 1, 1, 2, 2, 64),     # code layer: sigmoid, tiny planes
    (32, 192, 3, 2, 1, 16, 256),    # first layer: 3 input channels zero-padded to 32
    (192, 192, 3, 1, 1, 4, 512),    # wide rows: 128x1 tiles
    (96, 96, 3, 1, 1, 6, 320),      # 64x2 tiles, ragged rows
])
def test_conv_tensor_core_vs_direct(cuda, orc, Ci, Co, k, s, act, h, W):
    """tcgen05 TF32 implicit GEMM (NHWC) against the fp32 CUDA-core direct form (NCHW) on the same device.
    Tolerance: TF32 operands carry 10 explicit mantissa bits (activations truncated by the MMA, weights rounded
    when packed): per-output error ~ 2^-11 * sqrt(K) * rms(x) * rms(w) ~ 5e-4 here; asserted rms < 1.5e-3 and
    max < 1e-2 (x the gate magnitude), far below one quantiser step (~0.11)."""
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call
    rng = np.random.default_rng(31)
    halo = 2 if k == 3 else 0
    Hi, Wi = h + halo, W + halo
    ho, wo = (Hi - k) // s + 1, (Wi - k) // s + 1
    wl = [max(4, wo - 7 * (g % 5)) for g in range(16)]
    wl[3] = min(wo, 20)                                           # a band whose right-hand tiles are all invalid
    x = rng.standard_normal((16, Ci, Hi, Wi)).astype(np.float32)
    w = (rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)
    b = rng.standard_normal(Co).astype(np.float32)
    slope = rng.random(Co).astype(np.float32) * 0.5
    res = rng.standard_normal((16, Co, ho, wo)).astype(np.float32)
    mul = rng.standard_normal((16, Co, ho, wo)).astype(np.float32)
    dx, dw, db, ds, dm, dr = (T(a, cuda) for a in (x, w, b, slope, mul, res))
    # direct, NCHW
    d, Ho, Wo = _conv_desc(16, Ci, Hi, Wi, Co, k, s, act, wl, 1)
    y1 = torch.empty((16, Co, Ho, Wo), device=cuda)
    call("pcx_conv2d_fwd", C.byref(d), *(C.c_void_p(t.data_ptr()) for t in (dx, dw, db, ds, dm, dr, y1)), None)
    # tensor core, NHWC, output written into the interior of a larger (padded) plane
    d, Ho, Wo = _conv_desc(16, Ci, Hi, Wi, Co, k, s, act, wl, 0)
    d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = Ho + 2, Wo + 3, 1, 2
    xh, mh, rh = (t.permute(0, 2, 3, 1).contiguous() for t in (dx, dm, dr))
    y0 = torch.full((16, Ho + 2, Wo + 3, Co), -7.0, device=cuda)
    call("pcx_conv2d_fwd", C.byref(d), *(C.c_void_p(t.data_ptr()) for t in (xh, dw, db, ds, mh, rh, y0)), None)
    torch.cuda.synchronize()
    direct = N(y1)
    full = N(y0)
    tc = full[:, 1:1 + Ho, 2:2 + Wo, :].transpose(0, 3, 1, 2)
    border = full.copy()
    border[:, 1:1 + Ho, 2:2 + Wo, :] = -7.0
    assert (border == -7.0).all(), "cells outside the output window must stay untouched"
    for g in range(16):
        assert (tc[g, :, :, wl[g]:] == 0).all()
    err = np.abs(tc - direct)
    scale = max(1.0, float(np.abs(mul).max())) if act != 2 else float(np.abs(mul).max())
    assert np.sqrt((err ** 2).mean()) < 1.5e-3, "rms err %g" % np.sqrt((err ** 2).mean())
    assert err.max() < 1e-2 * scale, "max err %g" % err.max()


@pytest.mark.parametrize("act,Co,h,W", [(3, 192, 8, 128), (4, 192, 3, 72), (3, 96, 16, 520)])
def test_conv_square_input_is_the_conv_of_the_square(cuda, act, Co, h, W):
    """pcx_conv_desc.square_input (GDN / IGDN: the 1x1 GEMM over x^2, PseudoContextV2.py:186-216): squaring the activation stage
    in the kernel's shared-memory pipeline gives bit for bit the result of the same kernel reading a materialised x * x."""
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call
    rng = np.random.default_rng(77)
    Ci = Co
    wl = [max(4, W - 9 * (g % 4)) for g in range(16)]
    x = torch.from_numpy(rng.standard_normal((32, h, W, Ci)).astype(np.float32)).to(cuda)           # two images, NHWC
    w = torch.from_numpy((rng.random((Co, Ci, 1, 1)) * 0.02).astype(np.float32)).to(cuda)
    b = torch.from_numpy((1.0 + rng.random(Co)).astype(np.float32)).to(cuda)
    res = torch.from_numpy(rng.standard_normal((32, h, W, Co)).astype(np.float32)).to(cuda)
    P = lambda t: C.c_void_p(t.data_ptr())
    outs = []
    for fused in (0, 1):
        d, Ho, Wo = _conv_desc(32, Ci, h, W, Co, 1, 1, act, wl, 0)
        d.square_input = fused
        y = torch.full((32, h, W, Co), -3.0, device=cuda)
        call("pcx_conv2d_fwd", C.byref(d), P(x if fused else x * x), P(w), P(b), None, P(x), P(res), P(y), None)
        outs.append(y)
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1])
    # and the flag is refused where it has no meaning
    d, _, _ = _conv_desc(32, Ci, h, W, Co, 1, 1, 1, wl, 0)
    d.square_input = 1
    from pseudocylindrical_convolution_b200._lib import PcxError
    with pytest.raises(PcxError):
        call("pcx_conv2d_fwd", C.byref(d), P(x), P(w), P(b), P(b), None, None, P(outs[0]), None)


@pytest.mark.parametrize("k,h,W", [(3, 6, 256), (1, 5, 72), (3, 4, 640)])
def test_conv_fused_depth_to_space(cuda, k, h, W):
    """pcx_conv2d_fwd impl 3 (Dtow(2) folded into the store of a 192 -> 768 convolution: ResidualBlockUp conv1 / short_cut,
    model_zoo_v2.py:153-175) must equal impl 2 followed by pcx_dtow_nhwc BIT FOR BIT - same MMAs, same epilogue arithmetic,
    only the store addresses and the packed channel order differ.  CTA-pair kernel (k = 3) and single-CTA kernel (k = 1)."""
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200._lib import call
    rng = np.random.default_rng(77)
    Ci, Co, act = 192, 768, 1
    halo = 2 if k == 3 else 0
    Hi, Wi = h + halo, W + halo
    wl = [max(4, W - 9 * (g % 4)) for g in range(16)]
    wl[5] = min(W, 24)
    x = torch.from_numpy(rng.standard_normal((16, Hi, Wi, Ci)).astype(np.float32)).to(cuda)
    w = torch.from_numpy((rng.standard_normal((Co, Ci, k, k)) / np.sqrt(Ci * k * k)).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.standard_normal(Co).astype(np.float32)).to(cuda)
    slope = torch.from_numpy((rng.random(Co) * 0.5).astype(np.float32)).to(cuda)
    P = lambda t: C.c_void_p(t.data_ptr())

    def packed(fn):
        n = call(fn, None, None, Co, Ci, k, None)
        out = torch.empty(n, dtype=torch.float32, device=cuda)
        call(fn, P(w), P(out), Co, Ci, k, None)
        return out

    # two launches: conv into a plain tile, then the pixel shuffle into the interior of a padded plane
    d, Ho, Wo = _conv_desc(16, Ci, Hi, Wi, Co, k, 1, act, wl, 2)
    y = torch.empty((16, Ho, Wo, Co), device=cuda)
    call("pcx_conv2d_fwd", C.byref(d), P(x), P(packed("pcx_conv_pack_weights")), P(b), P(slope), None, None, P(y), None)
    rows, pitch, y0, x0 = 2 * Ho + 4, 2 * Wo + 5, 2, 3
    ref = torch.full((16, rows, pitch, Co // 4), -7.0, device=cuda)
    call("pcx_dtow_nhwc", P(y), P(ref), 16, Co // 4, Ho, Wo, Ho, Wo, 0, 0, rows, pitch, y0, x0, None)
    # one launch
    d, Ho, Wo = _conv_desc(16, Ci, Hi, Wi, Co, k, 1, act, wl, 3)
    d.out_rows, d.out_pitch, d.out_y0, d.out_x0 = rows, pitch, y0, x0
    got = torch.full((16, rows, pitch, Co // 4), -7.0, device=cuda)
    call("pcx_conv2d_fwd", C.byref(d), P(x), P(packed("pcx_conv_pack_weights_d2w")), P(b), P(slope), None, None, P(got), None)
    torch.cuda.synchronize()
    assert torch.equal(got, ref)
    assert float(ref[:, y0:y0 + 2 * Ho, x0:x0 + 2 * Wo].abs().max()) > 0.1
