"""CPU tests: the oracle's own invariants, the host-side geometry of libpcx against it, the range coder."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import W64, smooth_images


def test_set_weight_matches_survey_profile():
    from pseudocylindrical_convolution_b200.PCONV_operator import set_weight
    assert set_weight(16, True) == [float(v) for v in W64]


@pytest.mark.parametrize("W", [64, 128, 512, 1024, 2048, 8192])
def test_band_widths_product_vs_oracle(orc, W):
    from pseudocylindrical_convolution_b200 import _lib
    wl_o = orc.band_widths(W64, 16 * 4, W)
    w = (C.c_float * 16)(*[float(v) for v in W64])
    out = (C.c_int * 16)()
    _lib.call("pcx_band_widths", w, 16, 64, W, out)
    assert list(out) == list(wl_o)
    assert list(wl_o) == [int(v / 64 * W + 0.5) for v in W64]
    # valid fraction 836/1024 for every W that is a multiple of 64
    assert sum(wl_o) * 1024 == 836 * 16 * W


def test_band_widths_cosine_profile(orc):
    from pseudocylindrical_convolution_b200 import _lib
    from pseudocylindrical_convolution_b200.PCONV_operator import set_weight
    wt = set_weight(16, False)
    wt1 = [v / 64.0 for v in wt]              # total < 3*npart selects the cosine branch (math_cuda.cu:236-252)
    wl_o = orc.band_widths(wt1, 64, 256)
    w = (C.c_float * 16)(*wt1)
    out = (C.c_int * 16)()
    _lib.call("pcx_band_widths", w, 16, 64, 256, out)
    assert list(out) == list(wl_o)
    assert wl_o[7] == 256 and wl_o[8] == 256


def test_band_widths_rejects_bad_height():
    from pseudocylindrical_convolution_b200 import _lib
    w = (C.c_float * 16)(*[float(v) for v in W64])
    out = (C.c_int * 16)()
    with pytest.raises(_lib.PcxError):
        _lib.call("pcx_band_widths", w, 16, 65, 128, out)


def test_ctx_order_and_items_product_vs_oracle(orc):
    from pseudocylindrical_convolution_b200 import _lib
    for h, W, pad in [(2, 64, 2), (4, 128, 2), (1, 64, 1)]:
        wl = orc.band_widths(W64, 16 * h, W)
        g = orc.CtxGeom(wl, h, W, pad)
        Hf = 16 * h
        order = np.zeros(Hf * W, np.int32)
        start = np.zeros(Hf + W, np.int32)
        ip = C.POINTER(C.c_int)
        _lib.call("pcx_ctx_order", _lib.int_array(wl), 16, h, W, order.ctypes.data_as(ip), start.ctypes.data_as(ip))
        assert (start == g.start).all()
        n = start[-1]
        assert n == int(sum(wl)) * h
        assert (order[:n] == g.order[:n]).all()
        # every plane is sorted by row, cells are unique, plane index = row + col
        cells = order[:n]
        assert len(set(cells.tolist())) == n
        for p in range(Hf + W - 1):
            seg = cells[start[p]:start[p + 1]]
            assert ((seg // W + seg % W) == p).all()
            assert (np.diff(seg // W) > 0).all()
        pstart = np.zeros(Hf + W + pad, np.int32)
        cnt = _lib.call("pcx_ctx_pad_items", _lib.int_array(wl), 16, h, W, pad, g.band.reshape(-1).ctypes.data_as(ip),
                        g.col.reshape(-1).ctypes.data_as(ip), g.tw.reshape(-1).ctypes.data_as(C.POINTER(C.c_float)), None,
                        pstart.ctypes.data_as(ip))
        items = np.zeros((max(cnt, 1), 4), np.int32)
        _lib.call("pcx_ctx_pad_items", _lib.int_array(wl), 16, h, W, pad, g.band.reshape(-1).ctypes.data_as(ip),
                  g.col.reshape(-1).ctypes.data_as(ip), g.tw.reshape(-1).ctypes.data_as(C.POINTER(C.c_float)),
                  items.ctypes.data_as(ip), pstart.ctypes.data_as(ip))
        assert (pstart == g.pstart).all()
        assert (items[:cnt] == g.items[:cnt]).all()


def test_slice_uslice_oracle_properties(orc):
    """Size-independent properties: constant images survive slice and uslice exactly-ish, invalid columns are 0,
    full-width bands are (near) copies."""
    H, W = 64, 128
    wl = orc.band_widths(W64, H, W)
    x = np.full((1, 2, H, W), 0.625, np.float32)
    t = orc.sphere_slice(x, wl)
    assert t.shape == (16, 2, 4, W)
    for g in range(16):
        assert (t[g, :, :, wl[g]:] == 0).all()
        np.testing.assert_allclose(t[g, :, :, :wl[g]], 0.625, rtol=0, atol=2e-7)
    back = orc.sphere_uslice(t, wl)
    np.testing.assert_allclose(back, 0.625, rtol=0, atol=4e-7)
    img = smooth_images(1, 1, H, W)
    t = orc.sphere_slice(img, wl)
    full = [g for g in range(16) if wl[g] == W]
    assert full
    for g in full:   # column 0 carries the 1e-9 offset, the others are exact copies
        np.testing.assert_array_equal(t[g, 0, :, 1:], img[0, 0, g * 4:(g + 1) * 4, 1:])


def test_pad_oracle_properties(orc):
    h, W, pad = 4, 128, 2
    wl = orc.band_widths(W64, 16 * h, W)
    rng = np.random.default_rng(0)
    x = orc.pseudo_fill(rng.random((16, 3, h, W)).astype(np.float32), wl)
    y = orc.pseudo_pad(x, wl, pad)
    assert y.shape == (16, 3, h + 2 * pad, W + 2 * pad)
    for g in range(16):
        w = wl[g]
        np.testing.assert_array_equal(y[g, :, pad:pad + h, pad:pad + w], x[g, :, :, :w])
        np.testing.assert_array_equal(y[g, :, :, :pad], y[g, :, :, w:w + pad])                 # left wrap
        np.testing.assert_array_equal(y[g, :, :, pad + w:2 * pad + w], y[g, :, :, pad:2 * pad])  # right wrap
        assert (y[g, :, :, 2 * pad + w:] == 0).all()
    # seams between full-width bands are plain copies of the neighbour rows (weights collapse to 0/1 up to 1e-9)
    g = 7
    np.testing.assert_allclose(y[g, :, pad + h, pad:pad + W], x[g + 1, :, 0, :], rtol=0, atol=1e-6)
    # north pole: row -1 of band 0 is row 0 of band 0 shifted by half a turn
    np.testing.assert_allclose(y[0, :, pad - 1, pad:pad + wl[0]], np.roll(x[0, :, 0, :wl[0]], -(wl[0] // 2), axis=-1), rtol=0, atol=1e-6)


def test_quant_dquant_oracle_roundtrip(orc):
    h, W, Cc = 2, 64, 8
    wl = orc.band_widths(W64, 16 * h, W)
    theta = np.full((Cc, 8), np.log(1 / 9.0), np.float32)
    theta[:, 0] = 1 / 9.0
    steps = orc.quant_steps(theta)
    cen = orc.dquant_centres(theta)
    rng = np.random.default_rng(1)
    x = rng.random((16, Cc, h, W)).astype(np.float32)
    val, sym, count = orc.pseudo_quant(x, steps, wl)
    assert sym.min() >= 0 and sym.max() <= 7
    rec = orc.pseudo_dquant(sym, cen, wl)
    for g in range(16):
        assert (sym[g, :, :, wl[g]:] == 0).all() and (val[g, :, :, wl[g]:] == 0).all()
        np.testing.assert_allclose(rec[g, :, :, :wl[g]], val[g, :, :, :wl[g]], rtol=0, atol=1e-6)
        # nearest-centre rule
        d = np.abs(x[g, :, :, :wl[g], None] - cen[:, None, None, :])
        np.testing.assert_array_equal(np.argmin(d, -1), sym[g, :, :, :wl[g]].astype(int))
    assert count.sum() == -int(sum(wl)) * h * Cc


def test_dtow_oracle_inverse(orc):
    rng = np.random.default_rng(2)
    x = rng.random((3, 8, 4, 6)).astype(np.float32)
    y = orc.dtow(x, 2, True)
    assert y.shape == (3, 2, 8, 12)
    np.testing.assert_array_equal(orc.dtow(y, 2, False), x)
    import torch
    np.testing.assert_array_equal(y, torch.nn.functional.pixel_shuffle(torch.from_numpy(x), 2).numpy())


def test_gmm_table_oracle_properties(orc):
    rng = np.random.default_rng(3)
    n = 4000
    logit = rng.normal(size=(n, 3)).astype(np.float32)
    delta = (rng.normal(size=(n, 3)) * 1.5).astype(np.float32)          # negatives exercise the clamp
    mean = (rng.random((n, 3)) * 9 - 4.5).astype(np.float32)
    cdf, w, d = orc.gmm_table(logit, delta, mean)
    assert (cdf[:, 0] == 0).all() and (cdf[:, 8] == 65536).all()
    assert (np.diff(cdf, axis=1) > 0).all()                             # strictly increasing: every symbol codable
    np.testing.assert_allclose(w.sum(1), 1, atol=1e-6)
    assert (d >= 1e-6 * 0.999).all()
    # known-relation check of the reference's own __main__ (EntropyGmmTable.py:60-85): bin mass ~ exp(-nll)*65536
    lab = rng.integers(1, 7, size=n)
    nll = orc.gmm_nll(w, d, mean, (lab - 3.5).astype(np.float32))
    mass = cdf[np.arange(n), lab + 1] - cdf[np.arange(n), lab]
    np.testing.assert_allclose(np.exp(-nll) * 65536, mass, atol=9.0, rtol=3e-4)   # 2 roundings + fix-up shifts (<= 8)


def test_coder_bytes_identical_to_reference(ref_coder, tmp_path):
    """The product's host range coder must emit the reference coder's bitstream byte for byte."""
    import torch
    from pseudocylindrical_convolution_b200 import coder as mycoder
    rng = np.random.default_rng(4)
    n = 20000
    wgt = rng.integers(1, 3000, size=(n, 8)).astype(np.int64)
    wgt[rng.random(n) < 0.2] = [1, 1, 1, 60000, 1, 1, 1, 1]             # peaky rows
    cum = np.zeros((n, 9), np.int64)
    cum[:, 1:] = np.cumsum(wgt, 1)
    cum = cum * 65536 // cum[:, -1:]
    for j in range(8):
        cum[:, j + 1] = np.maximum(cum[:, j + 1], cum[:, j] + 1)
    cum = cum.astype(np.int32)
    sym = rng.integers(0, 8, size=n).astype(np.int32)
    t, s = torch.from_numpy(cum), torch.from_numpy(sym)
    mine = mycoder.coder(str(tmp_path / "mine.bin"))
    mine.start_encoder()
    for a in range(0, n, 777):                                           # ragged batches, like the wavefront
        b = min(n, a + 777)
        mine.encodes(t[a:b].contiguous(), 8, s[a:b].contiguous(), b - a)
    mine.end_encoder()
    data = open(tmp_path / "mine.bin", "rb").read()
    dec = mycoder.coder(str(tmp_path / "mine.bin"))
    dec.start_decoder()
    out = dec.decodes(t, 8, n)
    assert (out.numpy().astype(np.int32) == sym).all()
    golden = os.path.join(os.path.dirname(__file__), "golden", "coder_kat.npz")
    if os.path.exists(golden):
        kat = np.load(golden)
        k = mycoder.coder(str(tmp_path / "kat.bin"))
        k.start_encoder()
        k.encodes(torch.from_numpy(kat["cdf"]), 8, torch.from_numpy(kat["sym"]), len(kat["sym"]))
        k.end_encoder()
        assert open(tmp_path / "kat.bin", "rb").read() == kat["bytes"].tobytes()
    if ref_coder is None:
        pytest.skip("oracle/_ref/coder_ref.so not built; golden KAT checked only")
    r = ref_coder.coder(str(tmp_path / "ref.bin"))
    r.start_encoder()
    r.encodes(t, 8, s, n)
    r.end_encoder()
    assert open(tmp_path / "ref.bin", "rb").read() == data
    r2 = ref_coder.coder(str(tmp_path / "mine.bin"))
    r2.start_decoder()
    assert (r2.decodes(t, 8, n).numpy()[:n].astype(np.int32) == sym).all()


def test_coder_multibit_renormalisation_matches_reference(ref_coder, tmp_path):
    """The product coder shifts all agreeing / underflow bits in one go and divides by power-of-two totals with a shift;
    the reference shifts bit by bit and divides.  Width-1 symbols (16-bit bursts), long underflow runs (symbols straddling
    the midpoint) and non-power-of-two totals must give the same bytes, and both decoders must agree."""
    import torch
    from pseudocylindrical_convolution_b200 import coder as mycoder
    if ref_coder is None:
        pytest.skip("oracle/_ref/coder_ref.so not built")
    rng = np.random.default_rng(11)
    n = 30000
    cum = np.zeros((n, 9), np.int64)
    sym = np.zeros(n, np.int32)
    for i in range(n):
        kind = i % 4
        total = 65536 if kind != 3 else int(rng.integers(9, 70000))
        if kind == 0:                       # one dominant symbol, seven of width 1; code the rare ones often
            big = int(rng.integers(0, 8))
            w = np.ones(8, np.int64); w[big] = total - 7
            sym[i] = big if rng.random() < 0.5 else int(rng.integers(0, 8))
        elif kind == 1:                     # two halves meeting at the midpoint: underflow runs
            w = np.ones(8, np.int64); w[3] = total // 2 - 3; w[4] = total - 7 - w[3] + 1
            sym[i] = int(rng.choice([3, 4]))
        else:
            w = rng.integers(1, max(2, total // 8), size=8).astype(np.int64)
            w[-1] += max(0, total - w.sum())
            sym[i] = int(rng.integers(0, 8))
        c = np.concatenate([[0], np.cumsum(w)])
        cum[i] = c
    cum = cum.astype(np.int32)
    t, s_ = torch.from_numpy(cum), torch.from_numpy(sym)
    mine = mycoder.coder(str(tmp_path / "m.bin"))
    mine.start_encoder(); mine.encodes(t, 8, s_, n); mine.end_encoder()
    ref = ref_coder.coder(str(tmp_path / "r.bin"))
    ref.start_encoder(); ref.encodes(t, 8, s_, n); ref.end_encoder()
    assert open(tmp_path / "m.bin", "rb").read() == open(tmp_path / "r.bin", "rb").read()
    d = mycoder.coder(str(tmp_path / "r.bin"))
    d.start_decoder()
    assert (d.decodes(t, 8, n).numpy().astype(np.int32) == sym).all()


def test_coder_errors_are_loud(tmp_path):
    import torch
    from pseudocylindrical_convolution_b200 import coder as mycoder
    from pseudocylindrical_convolution_b200._lib import PcxError
    c = mycoder.coder(str(tmp_path / "x.bin"))
    c.start_encoder()
    bad = torch.tensor([[0, 10, 10, 30, 40, 50, 60, 70, 65536]], dtype=torch.int32)
    with pytest.raises(PcxError):
        c.encodes(bad, 8, torch.tensor([1], dtype=torch.int32), 1)      # zero-width symbol (ArithmeticCoder.cpp:45-46)
    empty = mycoder.coder(str(tmp_path / "none.bin"))
    with pytest.raises(PcxError):
        empty.start_decoder()


@pytest.mark.parametrize("ncode", [2, 5, 16, 32])
def test_coder_other_alphabet_sizes_match_reference(ref_coder, tmp_path, ncode):
    """The batched decoder counts scaled boundaries instead of dividing (pcx_coder.cpp); the codec only uses 8 symbols, the
    interface takes any ncode.  Bytes and symbols must match the reference coder for other alphabet sizes, power-of-two and
    arbitrary totals, ragged spans and a corrupted stream must fail loudly instead of decoding garbage silently."""
    import torch
    from pseudocylindrical_convolution_b200 import coder as mycoder
    from pseudocylindrical_convolution_b200._lib import PcxError
    if ref_coder is None:
        pytest.skip("oracle/_ref/coder_ref.so not built")
    rng = np.random.default_rng(100 + ncode)
    n = 6000
    cum = np.zeros((n, ncode + 1), np.int64)
    for i in range(n):
        total = 65536 if i % 3 else int(rng.integers(ncode + 1, 1 << 20))
        w = rng.integers(1, max(2, total // ncode), size=ncode).astype(np.int64)
        w[int(rng.integers(0, ncode))] += max(0, total - w.sum())
        cum[i, 1:] = np.cumsum(w)
    cum = cum.astype(np.int32)
    sym = rng.integers(0, ncode, size=n).astype(np.int32)
    t, s_ = torch.from_numpy(cum), torch.from_numpy(sym)
    mine = mycoder.coder(str(tmp_path / "m.bin"))
    mine.start_encoder()
    for a in range(0, n, 1234):
        b = min(n, a + 1234)
        mine.encodes(t[a:b].contiguous(), ncode, s_[a:b].contiguous(), b - a)
    mine.end_encoder()
    ref = ref_coder.coder(str(tmp_path / "r.bin"))
    ref.start_encoder(); ref.encodes(t, ncode, s_, n); ref.end_encoder()
    data = open(tmp_path / "m.bin", "rb").read()
    assert data == open(tmp_path / "r.bin", "rb").read()
    d = mycoder.coder(str(tmp_path / "r.bin"))
    d.start_decoder()
    out = np.concatenate([d.decodes(t[a:min(n, a + 777)].contiguous(), ncode, min(n, a + 777) - a).numpy() for a in range(0, n, 777)])
    assert (out.astype(np.int32) == sym).all()
    r2 = ref_coder.coder(str(tmp_path / "m.bin"))
    r2.start_decoder()
    assert (r2.decodes(t, ncode, n).numpy()[:n].astype(np.int32) == sym).all()
    # a decoder fed different tables than the encoder used either fails loudly or returns different symbols - never crashes
    other = cum.copy()
    other[:, 1:-1] = np.sort((other[:, 1:-1].astype(np.int64) * 7 // 8 + 1), axis=1)
    for j in range(ncode):
        other[:, j + 1] = np.maximum(other[:, j + 1], other[:, j] + 1)
    other[:, -1] = np.maximum(other[:, -1], cum[:, -1])
    d2 = mycoder.coder(str(tmp_path / "m.bin"))
    d2.start_decoder()
    try:
        got = d2.decodes(torch.from_numpy(other.astype(np.int32)), ncode, n).numpy().astype(np.int32)
        assert got.min() >= 0 and got.max() < ncode
    except PcxError:
        pass


def test_coder_rows16_decoder_matches_table_decoder(tmp_path):
    """pcx_coder_decodes_rows16 (the host half of the persistent decoder kernel, pcx_flow.cu): 16-byte rows = cum[1..7] as
    uint16 + a tag.  Same symbols as the int32-table decoder, stops at the first row whose tag is not the expected one, and
    writes (word_tag << 8 | symbol) words."""
    import ctypes as C
    import torch
    from pseudocylindrical_convolution_b200 import _lib
    from pseudocylindrical_convolution_b200 import coder as mycoder
    rng = np.random.default_rng(11)
    n = 5000
    wgt = rng.integers(1, 3000, size=(n, 8)).astype(np.int64)
    wgt[rng.random(n) < 0.3] = [1, 1, 60000, 1, 1, 1, 1, 1]
    cum = np.zeros((n, 9), np.int64)
    cum[:, 1:] = np.cumsum(wgt, 1)
    cum = cum * 65536 // cum[:, -1:]
    for j in range(8):
        cum[:, j + 1] = np.maximum(cum[:, j + 1], cum[:, j] + 1)
    cum[:, 8] = 65536
    for j in range(7, 0, -1):
        cum[:, j] = np.minimum(cum[:, j], cum[:, j + 1] - 1)
    assert (np.diff(cum, axis=1) > 0).all()
    sym = rng.integers(0, 8, size=n).astype(np.int32)
    enc = mycoder.coder(str(tmp_path / "s.bin"))
    enc.start_encoder()
    enc.encodes(torch.from_numpy(cum.astype(np.int32)), 8, torch.from_numpy(sym), n)
    enc.end_encoder()
    # aligned row buffer: 8 uint16 per row
    raw = np.zeros(n * 8 + 8, np.uint16)
    off = (-raw.ctypes.data // 2) % 8
    rows = raw[off:off + n * 8].reshape(n, 8)
    assert rows.ctypes.data % 16 == 0
    rows[:, :7] = cum[:, 1:8].astype(np.uint16)
    tag = 0x1234
    avail = 3210
    rows[:avail, 7] = tag                      # the "device" has produced the first `avail` rows only
    rows[avail:, 7] = tag - 1
    words = np.zeros(n, np.uint32)
    dec = mycoder.coder(str(tmp_path / "s.bin"))
    dec.start_decoder()
    done = C.c_int(0)
    pos = 0
    _lib.call("pcx_coder_decodes_rows16", dec._h, C.c_void_p(rows.ctypes.data), n, tag, 77, C.c_void_p(words.ctypes.data), C.byref(done))
    assert done.value == avail
    pos = done.value
    _lib.call("pcx_coder_decodes_rows16", dec._h, C.c_void_p(rows[pos:].ctypes.data), n - pos, tag, 77, C.c_void_p(words[pos:].ctypes.data), C.byref(done))
    assert done.value == 0                     # nothing new
    rows[avail:, 7] = tag
    _lib.call("pcx_coder_decodes_rows16", dec._h, C.c_void_p(rows[pos:].ctypes.data), n - pos, tag, 77, C.c_void_p(words[pos:].ctypes.data), C.byref(done))
    assert done.value == n - avail
    assert ((words & 0xff).astype(np.int32) == sym).all() and ((words >> 8) == 77).all()
    # a wrong table is reported like the table decoder reports it
    bad = mycoder.coder(str(tmp_path / "s.bin"))
    bad.start_decoder()
    rows2 = rows.copy()
    rows2[:, :7] = rows2[::-1, :7]
    raw2 = np.zeros(n * 8 + 8, np.uint16)
    off2 = (-raw2.ctypes.data // 2) % 8
    r2 = raw2[off2:off2 + n * 8].reshape(n, 8)
    r2[:] = rows2
    rc = _lib.load().pcx_coder_decodes_rows16(bad._h, C.c_void_p(r2.ctypes.data), n, tag, 77, C.c_void_p(words.ctypes.data), C.byref(done))
    assert rc in (0, -5) or rc < 0            # garbage tables either mis-decode silently or trip the bracket check; never crash
