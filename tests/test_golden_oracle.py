"""CPU: pins the oracle (oracle/pcx_oracle.c) to the committed golden vectors = outputs of the UNMODIFIED reference
extension run on a B200 (tools/make_golden.py, tests/golden/PROVENANCE.txt).

Bit-exact for every gather / integer / fixed-order fp32 path.  Where the GPU evaluates expf/erff (libdevice, MUFU.EX2)
and the oracle evaluates glibc's, the stated tolerance is: quantiser values 1 ulp-level (2e-7 relative), CDF entries
+-1 count."""
import os

import numpy as np
import pytest

import golden_cases as gc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    path = os.path.join(GOLD, "ref_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden file %s missing" % path)
    return np.load(path)


def _same_bits(got, want, what):
    got = np.ascontiguousarray(got, np.float32)
    want = np.ascontiguousarray(want, np.float32)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    bad = got.view(np.uint32) != want.view(np.uint32)
    assert not bad.any(), "%s: %d of %d values differ from the reference" % (what, int(bad.sum()), bad.size)


def test_provenance_is_recorded():
    assert os.path.exists(os.path.join(GOLD, "PROVENANCE.txt"))
    assert "PCONV_ref" in open(os.path.join(GOLD, "PROVENANCE.txt")).read()


def test_oracle_tile_ops_match_reference(orc):
    want = _gold("tiles")
    i = gc.inputs_tiles()
    wl = orc.band_widths(gc.W64, 32, 64)
    sl = orc.sphere_slice(i["erp"], wl)
    _same_bits(sl, want["slice"], "slice")
    _same_bits(orc.sphere_uslice(sl, wl), want["uslice"], "uslice")
    filled = orc.pseudo_fill(i["tiles"], wl)
    _same_bits(filled, want["fill"], "fill")
    for p in (1, 2):
        _same_bits(orc.pseudo_pad(filled, wl, p), want["pad%d" % p], "pad %d" % p)
    # PseudoFill asks the context for the widths of the PADDED extent (pseudo_fill_cuda.cu:12-25)
    wl_p = orc.band_widths(gc.W64, 16 * 6, 68)
    _same_bits(orc.pseudo_fill(want["pad2"], wl_p, pad=2, trim=1), want["fill_pad2_trim1"], "fill pad 2 trim 1")
    cf = orc.pseudo_fill(i["ctiles"], wl)
    for v in (1, 0):
        _same_bits(orc.pseudo_entropy_pad(cf, wl, 2, version=v), want["entropy_pad_v%d" % v], "entropy pad v%d" % v)
    d = orc.dtow(i["d2w"], 2, True)
    _same_bits(d, want["d2w"], "d2w")
    _same_bits(orc.dtow(d, 2, False), want["w2d"], "w2d")


def test_oracle_quantiser_matches_reference(orc):
    want = _gold("quant")
    i = gc.inputs_quant()
    wl = orc.band_widths(gc.W64, 32, 64)
    val, sym, _ = orc.pseudo_quant(i["x"], orc.quant_steps(i["theta"]), wl)
    assert np.array_equal(sym, want["sym"]), "symbols must be identical"
    np.testing.assert_allclose(val, want["val"], rtol=3e-7, atol=1e-7)
    dq = orc.pseudo_dquant(want["sym"], orc.dquant_centres(i["theta"]), wl)
    np.testing.assert_allclose(dq, want["dquant"], rtol=3e-7, atol=1e-7)


def test_oracle_gmm_matches_reference(orc):
    want = _gold("gmm")
    i = gc.inputs_gmm()
    cdf, w, dl = orc.gmm_table(i["logit"], i["delta"], i["mean"], form=0)
    _same_bits(dl, want["delta_clamped"], "delta clamp")
    np.testing.assert_allclose(w, want["softmax"], rtol=3e-6, atol=1e-8)
    for form, key in ((0, "cdf_batch"), (1, "cdf_plain")):
        got, _, _ = orc.gmm_table(i["logit"], i["delta"], i["mean"], form=form)
        ref = want[key].astype(np.int64)
        assert (ref[:, 0] == 0).all() and (ref[:, 8] == 65536).all() and (np.diff(ref, axis=1) > 0).all()
        diff = np.abs(got.astype(np.int64) - ref)
        assert diff.max() <= 1, key
        assert (diff > 0).mean() < 0.02, key
    nll = orc.gmm_nll(want["softmax"], want["delta_clamped"], i["mean"], i["label"].reshape(-1))
    np.testing.assert_allclose(np.exp(-nll.astype(np.float64)), np.exp(-want["nll"].astype(np.float64)), rtol=1e-5, atol=3e-7)


def test_oracle_wavefront_matches_reference_stream(orc, tmp_path):
    """The 12-layer masked context model stepped plane by plane: the extracted GMM parameters are bit-identical to the
    reference's at EVERY step (digest), the CDF rows agree to +-1 count, and coding the REFERENCE's tables with the
    product's host coder reproduces the reference bitstream byte for byte."""
    want = _gold("wavefront")
    i = gc.inputs_wavefront()
    G, h, W = i["G"], i["h"], i["W"]
    Hf = 16 * h
    wl = orc.band_widths(gc.W64, Hf, W)
    geom = orc.CtxGeom(wl, h, W, 2)
    NN = 3 * 16
    data = orc.pseudo_fill(i["sym"], wl)
    b_in = np.zeros((NN, G, h + 4, W + 4), np.float32)
    outs = [np.zeros((NN, 3 * G, h + 4, W + 4), np.float32) for _ in range(11)] + [np.zeros((NN, 3 * G, h, W), np.float32)]
    o_ext = np.zeros((3, 3, Hf, W), np.float32)
    o_lab = np.zeros((1, 1, Hf, W), np.float32)
    prev = np.zeros((1, 1, Hf, W), np.float32)

    def layer(k, x, s):
        w, b, a = i["layers"][k]
        orc.ctx_pad_step(x, geom, G, s - 1 if k == 0 else s)
        orc.ctx_conv_step(x, w, b, a, outs[k], geom, G, 1, 2, 0 if k == 11 else 2, 5 if k == 0 else 6, s)
        return outs[k]

    row = 0
    worst = 0
    for s in range(geom.nsteps(G)):
        orc.dinput_step(prev.reshape(-1), b_in, geom, G, 1, 2, -3.5, 3, s)
        x = layer(0, b_in, s)
        for blk in range(5):
            y = layer(2 + 2 * blk, layer(1 + 2 * blk, x, s), s)
            x = orc.ctx_add_step(y, x, geom, G, 2, s)
        n = orc.dextract_step(layer(11, x, s), o_ext, geom, G, s, True)
        assert n == int(want["counts"][s]), "symbol count at step %d" % s
        n_lab = orc.dextract_step(data, o_lab, geom, G, s, False)
        prev = np.zeros((1, 1, Hf, W), np.float32)
        prev.reshape(-1)[:n_lab] = o_lab.reshape(-1)[:n_lab]
        if n == 0:
            continue
        z = o_ext.reshape(3, -1)[:, :n * 3]
        assert gc._fold(z) == want["digests"][s], "GMM parameters differ from the reference at step %d" % s
        assert np.array_equal(o_lab.reshape(-1)[:n].astype(np.uint8), want["labels"][row:row + n])
        cdf, _, _ = orc.gmm_table(z[0].reshape(n, 3), z[1].reshape(n, 3), z[2].reshape(n, 3))
        worst = max(worst, int(np.abs(cdf.astype(np.int64) - want["tables"][row:row + n]).max()))
        row += n
    assert row == len(want["tables"]) == int(wl.sum()) * h * G
    assert worst <= 1

    from pseudocylindrical_convolution_b200 import coder
    import torch
    enc = coder.coder(str(tmp_path / "mine.bin"))
    enc.start_encoder()
    row = 0
    for n in want["counts"]:
        n = int(n)
        if n:
            enc.encodes(torch.from_numpy(want["tables"][row:row + n].copy()), 8,
                        torch.from_numpy(want["labels"][row:row + n].astype(np.int32)), n)
            row += n
    enc.end_encoder()
    assert np.array_equal(np.frombuffer(open(tmp_path / "mine.bin", "rb").read(), np.uint8), want["bitstream"])
    # and decoding the reference's bitstream with the reference's tables returns the reference's symbols
    dec = coder.coder(str(tmp_path / "mine.bin"))
    dec.start_decoder()
    row = 0
    for n in want["counts"]:
        n = int(n)
        if n:
            sym = dec.decodes(torch.from_numpy(want["tables"][row:row + n].copy()), 8, n)
            assert np.array_equal(sym.numpy()[:n].astype(np.uint8), want["labels"][row:row + n])
            row += n
