"""BASELINE configs[0]: pseudo_codec model-idx 3 (--ssim -> prefix '4_56', valid_dim 56) encode + decode of one synthetic
512x1024 ERP image with seeded random-init weights, through the product's PseudoEncoder / PseudoDecoder (= the reference's
classes, same state-dict keys, strict load).

  * the decoder recovers the encoder's symbol tensor bit for bit from the bitstream (what decodability needs: both sides run
    the same CUDA-core entropy path, whose arithmetic is pinned against the reference in test_golden_gpu.py);
  * the tensor-core (TF32) transforms agree with the fp32 exact-order path and with a plain PyTorch fp32 evaluation of the
    reference's layer sequence within the stated tolerances; symbol flips near bin edges are counted, not hidden."""
import os

import numpy as np
import pytest

from conftest import smooth_images

pytestmark = pytest.mark.gpu
H, W, VD, PREX = 512, 1024, 56, "4_56"


@pytest.fixture(scope="module")
def codec(cuda, tmp_path_factory):
    import torch
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    d = str(tmp_path_factory.mktemp("demo_ssim"))
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, PREX, VD, 0, seed=0)
    enc = pc.PseudoEncoder(VD, 0).to(cuda)
    dec = pc.PseudoDecoder(VD, 0).to(cuda)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")          # strict: the key sets are the reference's
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    x = torch.from_numpy(smooth_images(1, 3, H, W, seed=1234)).to(cuda)
    return enc, dec, x, d


def test_state_dict_contract(codec):
    enc, dec, _, _ = codec
    es, ds = enc.state_dict(), dec.state_dict()
    assert len(es) == 186 and sum(v.numel() for v in es.values()) == 7237914
    assert len(ds) == 191 and sum(v.numel() for v in ds.values()) == 11275302
    assert tuple(es["ent.net.0.conv.weight"].shape) == (3, 42, 14, 5, 5)
    assert tuple(ds["decoder.net.3.conv1.weight"].shape) == (768, 192, 3, 3)


def test_encode_decode_roundtrip_is_lossless_on_symbols(codec, tmp_path):
    import torch
    enc, dec, x, _ = codec
    sym = enc.symbols(x)
    assert tuple(sym.shape) == (16, VD // 4, H // 128, W // 8)
    s = sym.cpu().numpy()
    assert set(np.unique(s)) <= set(range(8)) and len(np.unique(s)) >= 4
    path = str(tmp_path / "img.bin")
    enc(x, path)
    nbytes = os.path.getsize(path)
    bpp = nbytes * 8 / (H * W)
    assert 0.01 < bpp < 3.0 * VD / 256 * 0.8164 + 0.1, bpp       # at most 3 bits per 8-level symbol
    rec = dec(path, H, W)
    assert tuple(rec.shape) == (1, 3, H, W) and torch.isfinite(rec).all()
    # the decoder's entropy stage alone returns the symbol tensor
    dec.ent.start(path)
    got = dec.ent(H // 128, W // 8)
    assert torch.equal(got, sym), "decoded symbols differ from the encoded ones"
    # decoding twice gives the same bytes -> same reconstruction (deterministic kernels)
    rec2 = dec(path, H, W)
    assert torch.equal(rec, rec2)
    print("bpp %.4f, %d bytes" % (bpp, nbytes))


def _torch_reference_encoder(enc, x):
    """The reference's layer sequence (model_zoo_v2.py:95-151) evaluated with plain PyTorch fp32 convolutions (TF32 off) and
    this package's NCHW pad / fill operators (bit-exact to the reference's, test_golden_gpu.py)."""
    import torch
    import torch.nn.functional as F
    e = enc.encoder

    def fill(m, t):
        return m.trim(t.contiguous())

    def rb(m, t):
        y = F.prelu(F.conv2d(m.pad(t), m.conv1.weight, m.conv1.bias), m.relu1.weight)      # no fill in between (reference :49-53)
        y = F.prelu(F.conv2d(y, m.conv2.weight, m.conv2.bias), m.relu2.weight)
        y = F.conv2d(y, m.conv3.weight, m.conv3.bias)
        return fill(m, t + y)

    def rbv2(m, t):
        y = F.prelu(F.conv2d(m.pad(t), m.conv1.weight, m.conv1.bias), m.relu1.weight)
        y = F.prelu(F.conv2d(y, m.conv2.weight, m.conv2.bias), m.relu2.weight)
        return fill(m, t + y)

    def gdn(m, t):
        be, ge = m._effective()
        NN, ch, h, Wd = t.shape
        wl = m.ctx[0].op[0].widths(h, Wd)
        mask = torch.zeros((NN, 1, 1, Wd), device=t.device)
        for n in range(NN):
            mask[n, 0, 0, :wl[n % 16]] = 1
        t = t * mask
        norm = torch.sqrt(F.conv2d(t * t, ge.view(ch, ch, 1, 1), be))
        norm = norm * mask + 1 - mask
        return t / norm

    def down(m, t):
        y = fill(m, F.prelu(F.conv2d(m.pad1(t), m.conv1.weight, m.conv1.bias, stride=2), m.relu1.weight))
        y = gdn(m.relu2, fill(m, F.conv2d(m.pad2(y), m.conv2.weight, m.conv2.bias)))
        s = F.conv2d(t, m.short_cut.weight, m.short_cut.bias, stride=2)
        return fill(m, s + y)

    def att(m, t):
        a = t
        tr = t
        for i in range(3):
            tr = rb(m.trunk[i], tr)
            a = rb(m.attention[i], a)
        a = torch.sigmoid(F.conv2d(a, m.attention[3].weight, m.attention[3].bias))
        return fill(m, t + tr * a)

    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            t = enc.slice(x)
            t = down(e.net[0], t); t = rbv2(e.net[1], t); t = down(e.net[2], t); t = att(e.net[3], t)
            t = rbv2(e.net[4], t); t = down(e.net[5], t); t = rbv2(e.net[6], t)
            m = e.net[7]
            t = fill(m, F.conv2d(m.pad(t), m.conv.weight, m.conv.bias, stride=2))
            t = att(e.net[8], t)
            return e.trim(torch.sigmoid(F.conv2d(t, e.net[9].weight, e.net[9].bias)).contiguous())
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def test_tensor_core_transforms_vs_fp32_paths(codec):
    """Analysis transform: tcgen05 TF32 path vs (a) the NCHW fp32 exact-order kernels, (b) plain PyTorch fp32.
    Tolerance on the sigmoid code (range [0,1], quantiser step ~0.12): max |diff| < 2e-2, rms < 2e-3; symbol flips < 2 %."""
    import torch
    from pseudocylindrical_convolution_b200 import config
    enc, dec, x, _ = codec
    lat_tc = enc.latent(x).clone()
    sym_tc = enc.quant(lat_tc)[1].clone()
    config.CONV_IMPL = 1
    try:
        lat_f = enc.latent(x).clone()
    finally:
        config.CONV_IMPL = 0
    sym_f = enc.quant(lat_f)[1].clone()
    ref = _torch_reference_encoder(enc, x)
    q_ref = enc.quant(ref)[1].clone()
    for name, a, b in (("tc vs fp32 direct", lat_tc, lat_f), ("fp32 direct vs torch fp32", lat_f, ref), ("tc vs torch fp32", lat_tc, ref)):
        d = (a - b).abs()
        print("%s: max %.3e rms %.3e" % (name, float(d.max()), float((d ** 2).mean().sqrt())))
    d = (lat_f - ref).abs()
    assert float(d.max()) < 2e-3, "the fp32 exact-order path must agree with PyTorch fp32 to rounding noise"
    d = (lat_tc - ref).abs()
    assert float(d.max()) < 2e-2 and float((d ** 2).mean().sqrt()) < 2e-3
    flips = float((sym_tc != q_ref).float().mean())
    print("symbol flips tc vs torch-fp32: %.4f %%, fp32-direct vs torch-fp32: %.4f %%" % (100 * flips, 100 * float((sym_f != q_ref).float().mean())))
    assert flips < 0.02
    assert float((sym_f != q_ref).float().mean()) < 0.002
    code_tc = lat_tc
    # invalid columns stay exactly zero on both paths
    wl = enc.ctx.op[0].widths(code_tc.shape[2], code_tc.shape[3])
    for g in range(16):
        assert float(code_tc[g, :, :, wl[g]:].abs().max()) == 0.0 if wl[g] < code_tc.shape[3] else True


def test_synthesis_transform_tc_vs_fp32(codec):
    """Synthesis transform on the same symbols: reconstructions of the TF32 and fp32 exact-order paths agree to
    max |diff| < 3e-2 (image range [0,1]) and PSNR between them > 50 dB."""
    import math
    import torch
    from pseudocylindrical_convolution_b200 import config
    enc, dec, x, _ = codec
    sym = enc.symbols(x)
    rec_tc = dec.reconstruct(sym).clone()
    config.CONV_IMPL = 1
    try:
        rec_f = dec.reconstruct(sym).clone()
    finally:
        config.CONV_IMPL = 0
    d = (rec_tc - rec_f).abs()
    mse = float((d ** 2).mean())
    print("recon tc vs fp32: max %.3e, PSNR between %.1f dB" % (float(d.max()), 10 * math.log10(1.0 / max(mse, 1e-20))))
    assert torch.isfinite(rec_tc).all()
    assert float(d.max()) < 3e-2 and mse < 1e-5


def test_native_wavefront_engine_equals_operator_loop(codec, tmp_path):
    """The native engine (one call for the whole serial loop) and the operator-by-operator loop (the reference's shape)
    must write byte-identical bitstreams and decode identical symbols."""
    import torch
    from pseudocylindrical_convolution_b200 import config
    enc, dec, x, _ = codec
    sym = enc.symbols(x)
    files = {}
    for impl in (1, 0):
        config.WAVE_IMPL = impl
        try:
            path = str(tmp_path / ("wave%d.bin" % impl))
            enc.ent.start(path)
            enc.ent(sym.clone())
            files[impl] = open(path, "rb").read()
            dec.ent.start(path)
            got = dec.ent(H // 128, W // 8)
            assert torch.equal(got, sym), "impl %d: decoded symbols differ" % impl
        finally:
            config.WAVE_IMPL = 1
    assert files[0] == files[1] and len(files[0]) > 1000


def test_batched_codec_equals_single_image_streams(codec, tmp_path):
    """Three images through the wavefront TOGETHER (nimg = 3 in every launch) give the same three bitstreams as coding
    them one by one, and batched decoding returns every image's symbols and reconstruction."""
    import torch
    enc, dec, x, _ = codec
    xs = torch.cat([x, torch.from_numpy(smooth_images(2, 3, H, W, seed=77)).to(x.device)]).contiguous()
    singles = []
    for i in range(3):
        p = str(tmp_path / ("single%d.bin" % i))
        enc(xs[i:i + 1].contiguous(), p)
        singles.append(open(p, "rb").read())
    names = [str(tmp_path / ("batch%d.bin" % i)) for i in range(3)]
    enc.encode_batch(xs, names)
    for i in range(3):
        assert open(names[i], "rb").read() == singles[i], "image %d: batched bitstream differs" % i
    sym = enc.symbols(xs)
    got = dec.ent.decode_batch(H // 128, W // 8, names)
    assert torch.equal(got, sym)
    rec = dec.decode_batch(names, H, W)
    assert tuple(rec.shape) == (3, 3, H, W)
    one = dec(names[1], H, W)
    assert float((rec[1:2] - one).abs().max()) < 1e-6


def test_one_shot_encoder_equals_stepwise_engine(codec, tmp_path):
    """pcx_wave_encode_full evaluates every context-model layer over the whole symbol tensor in one launch and emits the CDF rows
    in coding order; the stepwise engine runs the 204-step loop.  Same per-scalar arithmetic -> byte-identical bitstreams, for one
    image, for a batch, and when the CDF stream is cut into many chunks."""
    import torch
    from pseudocylindrical_convolution_b200 import config
    enc, dec, x, _ = codec
    xs = torch.cat([x, torch.from_numpy(smooth_images(2, 3, H, W, seed=99)).to(x.device)]).contiguous()
    sym = enc.symbols(xs)
    streams = {}
    try:
        for full, chunk in ((0, 1 << 17), (1, 1 << 17), (1, 3000)):
            config.WAVE_ENCODE_FULL, config.WAVE_CHUNK_ROWS = full, chunk
            p1 = str(tmp_path / ("one_%d_%d.bin" % (full, chunk)))
            enc.ent.start(p1)
            enc.ent(sym[:16].clone())
            names = [str(tmp_path / ("b%d_%d_%d.bin" % (i, full, chunk))) for i in range(3)]
            enc.ent.encode_batch(sym.clone(), names)
            streams[(full, chunk)] = [open(p1, "rb").read()] + [open(n, "rb").read() for n in names]
    finally:
        config.WAVE_ENCODE_FULL, config.WAVE_CHUNK_ROWS = 1, 1 << 17
    ref = streams[(0, 1 << 17)]
    assert len(ref[0]) > 1000 and ref[0] == ref[1]
    for key, got in streams.items():
        assert got == ref, "bitstreams of %s differ from the stepwise engine" % (key,)
    got = dec.ent.decode_batch(H // 128, W // 8, names)
    assert torch.equal(got, sym)


def test_fused_step_decoder_equals_operator_sequence(codec, tmp_path):
    """The three decoder engines - 2: one persistent dataflow kernel per decode (pcx_flow.cu: write-once scratch polled scalar by
    scalar, 16-byte CDF rows and symbol words through mapped pinned memory, per-image pipelines), 1: one cooperative launch per
    wavefront step with grid barriers, 0: the launch-per-operator sequence - must decode the same symbols, for one image and for
    a batch, from bitstreams written by the one-shot encoder."""
    import torch
    from pseudocylindrical_convolution_b200 import _lib
    enc, dec, x, _ = codec
    xs = torch.cat([x, torch.from_numpy(smooth_images(2, 3, H, W, seed=5)).to(x.device)]).contiguous()
    sym = enc.symbols(xs)
    names = [str(tmp_path / ("f%d.bin" % i)) for i in range(3)]
    enc.ent.encode_batch(sym.clone(), names)
    lib = _lib.load()
    try:
        for fused in (2, 1, 0):
            lib.pcx_wave_set_fused(fused)
            n0 = lib.pcx_launch_count()
            got = dec.ent.decode_batch(H // 128, W // 8, names)
            launches = lib.pcx_launch_count() - n0
            assert torch.equal(got, sym), "fused=%d: batch decode differs" % fused
            dec.ent.start(names[1])
            one = dec.ent(H // 128, W // 8)
            assert torch.equal(one, sym[16:32]), "fused=%d: single-image decode differs" % fused
            if fused == 1:
                assert launches <= 204 + 8, launches          # one launch per step (+ the final DInput2 / fill)
            if fused == 2:
                assert launches <= 8, launches                # sentinel fill, cell table, THE kernel (+ fill)
        # engine 2 with fewer host decoder threads than images (one thread serves several pipelines) and repeated calls
        lib.pcx_wave_set_fused(2)
        for nthreads in (1, 2):
            prev = lib.pcx_flow_set_threads(nthreads)
            try:
                assert torch.equal(dec.ent.decode_batch(H // 128, W // 8, names), sym), "threads=%d" % nthreads
            finally:
                lib.pcx_flow_set_threads(prev)
        # both forms of the dataflow kernel (256 threads x 2 blocks per SM, 128 x 4) on the same streams, whatever the automatic choice
        for form in ("0", "1"):
            os.environ["PCX_FLOW_FORM"] = form
            try:
                assert torch.equal(dec.ent.decode_batch(H // 128, W // 8, names), sym), "kernel form %s" % form
                dec.ent.start(names[2])
                assert torch.equal(dec.ent(H // 128, W // 8), sym[32:48]), "kernel form %s, single image" % form
            finally:
                del os.environ["PCX_FLOW_FORM"]
    finally:
        lib.pcx_wave_set_fused(2)


def test_flow_decoder_reports_a_corrupt_stream(codec, tmp_path):
    """A truncated / foreign bitstream must end the persistent decoder kernel with an error (or, if the garbage happens to stay
    inside every bracket, with wrong symbols) - never with a hang: the host sets the abort word, every device wait checks it."""
    import torch
    from pseudocylindrical_convolution_b200 import _lib
    enc, dec, x, _ = codec
    sym = enc.symbols(x)
    good = str(tmp_path / "g.bin")
    enc.ent.encode_batch(sym.clone(), [good])
    data = open(good, "rb").read()
    bad = str(tmp_path / "b.bin")
    rng = np.random.default_rng(3)
    open(bad, "wb").write(bytes(rng.integers(0, 256, size=len(data) // 2, dtype=np.uint8)))
    lib = _lib.load()
    lib.pcx_wave_set_fused(2)
    try:
        got = dec.ent.decode_batch(H // 128, W // 8, [bad])
        assert not torch.equal(got, sym)
    except _lib.PcxError as e:
        assert "coder" in str(e) or "decoder" in str(e)
    # the engine is usable afterwards
    assert torch.equal(dec.ent.decode_batch(H // 128, W // 8, [good]), sym)


@pytest.mark.parametrize("vd,prex,Hs,Ws", [(112, "5_112", 512, 1024), (192, "8_192", 256, 512)])
def test_other_model_sizes_round_trip(cuda, tmp_path, vd, prex, Hs, Ws):
    """model-idx 4..8 use 112 / 192 code channels (28 / 48 channel groups in the context model): every engine must handle them
    (larger weight rows in shared memory, longer chains) - symbols decode losslessly in all decoder modes and the one-shot
    and stepwise encoders agree."""
    import torch
    from pseudocylindrical_convolution_b200 import _lib, config, pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    d = str(tmp_path)
    p_enc, p_dec, p_ent = synthesize_checkpoints(d, prex, vd, 0, seed=3)
    enc = pc.PseudoEncoder(vd, 0).to(cuda)
    dec = pc.PseudoDecoder(vd, 0).to(cuda)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    x = torch.from_numpy(smooth_images(2, 3, Hs, Ws, seed=21)).to(cuda)
    sym = enc.symbols(x)
    assert tuple(sym.shape) == (32, vd // 4, Hs // 128, Ws // 8)
    streams = {}
    try:
        for full in (1, 0):
            config.WAVE_ENCODE_FULL = full
            names = [str(tmp_path / ("m%d_%d.bin" % (full, i))) for i in range(2)]
            enc.ent.encode_batch(sym.clone(), names)
            streams[full] = [open(n, "rb").read() for n in names]
    finally:
        config.WAVE_ENCODE_FULL = 1
    assert streams[0] == streams[1] and len(streams[0][0]) > 100
    # What the coder sees is the PseudoFill'ed tensor (pseudo_codec.py:99).  For widths that are not a multiple of 1024 the band
    # widths at the code scale do not double exactly into those of the context scale (2*round(7.5) != 15), so
    # a column the quantiser filled can lie outside the context geometry - the reference drops it the same way.
    want = enc.ent.fill(sym.clone())
    lib = _lib.load()
    try:
        for fused in (2, 1, 0):
            lib.pcx_wave_set_fused(fused)
            got = dec.ent.decode_batch(Hs // 128, Ws // 8, names)
            assert torch.equal(got, want), "vd %d fused %d" % (vd, fused)
    finally:
        lib.pcx_wave_set_fused(2)
    rec = dec.decode_batch(names, Hs, Ws)
    assert tuple(rec.shape) == (2, 3, Hs, Ws) and torch.isfinite(rec).all()


def test_graph_replay_equals_eager_and_tracks_parameters(codec):
    """The channels-last transforms are replayed from a CUDA graph from the third call with the same problem
    (transforms_nhwc._run_graphed).  Replays must be bit-identical to the eager pass, a new input must flow through the
    graph's static buffer, and an in-place parameter update or a new shape must drop the capture (never a stale result)."""
    import torch
    from pseudocylindrical_convolution_b200 import config
    enc, dec, x, _ = codec
    x2 = torch.from_numpy(smooth_images(1, 3, H, W, seed=77)).to(x.device)

    def eager(t):
        config.CUDA_GRAPHS = False
        try:
            return enc.latent(t).clone()
        finally:
            config.CUDA_GRAPHS = True

    ref1, ref2 = eager(x), eager(x2)
    assert not torch.equal(ref1, ref2)
    enc.encoder.__dict__.pop("_pcx_graph", None)
    outs = [enc.latent(x).clone() for _ in range(4)]                  # eager, eager + capture, replay, replay
    st = enc.encoder.__dict__["_pcx_graph"]
    assert st["graph"] is not None and not st["failed"], "the transform should have been captured"
    for o in outs:
        assert torch.equal(o, ref1)
    assert torch.equal(enc.latent(x2), ref2)                          # replay on a different image
    # in-place parameter update -> the version stamp changes -> eager pass with the new weights, then a new capture
    wt = enc.encoder.net[9].weight
    saved = wt.detach().clone()
    try:
        with torch.no_grad():
            wt.mul_(0.5)
        ref3 = eager(x)
        assert not torch.equal(ref3, ref1)
        for _ in range(3):
            assert torch.equal(enc.latent(x), ref3)
    finally:
        with torch.no_grad():
            wt.copy_(saved)
    # a different problem size in between drops the capture; coming back re-captures with identical results
    xs = torch.from_numpy(smooth_images(1, 3, 256, 512, seed=5)).to(x.device)
    small = [enc.latent(xs).clone() for _ in range(3)]
    assert torch.equal(small[0], small[2])
    for _ in range(3):
        assert torch.equal(enc.latent(x), ref1)
    # synthesis side
    sym = enc.symbols(x)
    config.CUDA_GRAPHS = False
    try:
        rec0 = dec.reconstruct(sym).clone()
    finally:
        config.CUDA_GRAPHS = True
    for _ in range(4):
        assert torch.equal(dec.reconstruct(sym), rec0)


def test_one_shot_encoder_variants_emit_identical_streams(codec, tmp_path):
    """The one-shot entropy encoder has three interchangeable execution forms: wavefront slabs (the host codes slab k while the
    device computes slab k + 1), the shared-memory context convolution with its channel-group pairs dealt to 1..7 blocks, and the
    L1-resident tiled convolution.  Every combination must write the same bytes as the default, and the default decodes."""
    import torch
    from pseudocylindrical_convolution_b200 import _lib
    enc, dec, x, _ = codec
    lib = _lib.load()
    sym = enc.symbols(x)

    def encode(tag, **opts):
        prev = {k: lib.pcx_wave_set_option(k.encode(), v) for k, v in opts.items()}
        try:
            path = str(tmp_path / (tag + ".bin"))
            enc.ent.encode_batch(sym, [path])
            return open(path, "rb").read()
        finally:
            for k, v in prev.items():
                lib.pcx_wave_set_option(k.encode(), v)

    ref = encode("default")
    assert len(ref) > 1000
    for tag, opts in (("slabs2", dict(slabs=2)), ("slabs3", dict(slabs=3)), ("slabs5_t7", dict(slabs=5, tsplit=7)), ("t1", dict(tsplit=1)),
                      ("t3", dict(tsplit=3)), ("l1", dict(smem=0)), ("l1_slabs4", dict(smem=0, slabs=4))):
        assert encode(tag, **opts) == ref, tag
    assert lib.pcx_wave_set_option(b"nonsense", 1) < 0
    got = dec.ent.decode_batch(H // 128, W // 8, [str(tmp_path / "default.bin")])
    assert torch.equal(got, sym)


def test_refresh_after_data_edit(codec):
    """Parameter edits through `.data` do not bump Tensor._version, which the packed-weight / GDN / CUDA-graph caches key on:
    `refresh()` (also run by load_state_dict and .to()) must make the next pass use the new values."""
    import torch
    enc, dec, x, _ = codec
    for _ in range(3):
        ref = enc.latent(x).clone()                                # eager, capture, replay
    wt = enc.encoder.net[9].weight
    saved = wt.data.clone()
    try:
        wt.data.mul_(0.5)                                          # invisible to the version stamps
        enc.encoder.refresh()
        got = enc.latent(x).clone()
        assert not torch.equal(got, ref), "stale packed weights / graph after refresh()"
        sd = {k: v.clone() for k, v in enc.encoder.state_dict().items()}
        sd["net.9.weight"] = saved
        enc.encoder.load_state_dict(sd)                            # the post hook refreshes
        assert torch.equal(enc.latent(x), ref)
    finally:
        wt.data.copy_(saved)
        enc.encoder.refresh()
