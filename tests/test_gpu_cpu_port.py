"""The product's transforms against the CPU port of the reference codec (oracle/cpu_codec.py) - the INDEPENDENT oracle of the
layer topology: analysis (EncoderV2) and, above all, synthesis (DecoderV2 / ResidualBlockUp / IGDN wiring, model_zoo_v2.py:
153-211), which round 1 only compared with itself.  The port evaluates the reference's graph with the C restatement of every
custom operator (pinned to the unmodified reference extension by tests/test_golden_oracle.py) and torch-CPU fp32 convolutions.
Checkpoints come from the port's seeded generator and are strict-loaded by the product: the key sets must be the reference's.

Tolerances (stated): the NCHW fp32 exact-order path agrees with the port to float rounding (max |diff| < 2e-4 on
values in [0, 1]); the tcgen05 TF32 path within max < 2e-2 / rms < 2e-3 on the sigmoid code and max < 5e-2, PSNR > 40 dB on the
reconstruction.  Also: round trips at the sizes the reference cannot run (1024x2048, 2048x4096: auto slabs in the one-shot
encoder, > 2^24 offsets, 396 / 780 wavefront steps)."""
import os

import numpy as np
import pytest

from conftest import smooth_images

pytestmark = pytest.mark.gpu
VD, PREX = 56, "4_56"


@pytest.fixture(scope="module")
def both(cuda, tmp_path_factory):
    import torch  # noqa: F401
    from oracle import cpu_codec as cc
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    d = str(tmp_path_factory.mktemp("cpu_port"))
    (p_enc, p_dec, p_ent), sd = cc.save_checkpoints(d, PREX, VD, seed=5)
    enc = pc.PseudoEncoder(VD, 0).to(cuda)
    dec = pc.PseudoDecoder(VD, 0).to(cuda)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    return enc, dec, cc.CpuCodec(sd, VD), d


def _with_impl(impl, fn):
    from pseudocylindrical_convolution_b200 import config
    old = config.CONV_IMPL
    config.CONV_IMPL = impl
    try:
        return fn()
    finally:
        config.CONV_IMPL = old


def test_analysis_transform_vs_cpu_port(both, cuda):
    import torch
    enc, dec, cpu, _ = both
    x = smooth_images(1, 3, 256, 512, seed=3)
    want = cpu.analysis(x)
    xt = torch.from_numpy(x).to(cuda)
    fp32 = _with_impl(1, lambda: enc.latent(xt).cpu().numpy())
    tc = _with_impl(0, lambda: enc.latent(xt).cpu().numpy())
    assert want.shape == fp32.shape == tc.shape == (16, 192, 1, 32)
    assert np.abs(fp32 - want).max() < 2e-4, np.abs(fp32 - want).max()
    assert np.abs(tc - want).max() < 2e-2 and np.sqrt(np.mean((tc - want) ** 2)) < 2e-3
    # symbols: the fp32 path may differ from the port only where the code sits within float noise of a bin edge
    sym_cpu = cpu.symbols(x)
    sym_fp32 = _with_impl(1, lambda: enc.symbols(xt).cpu().numpy())
    assert (sym_cpu != sym_fp32).mean() < 1e-3


def test_synthesis_transform_vs_cpu_port(both, cuda):
    """DecoderV2 topology: product (both convolution paths) vs the port, from the SAME symbol tensor."""
    import torch
    enc, dec, cpu, _ = both
    x = smooth_images(1, 3, 256, 512, seed=4)
    sym = cpu.symbols(x)                                  # (16, 14, 2, 64) symbols 0..7 from the port's own encoder
    assert len(np.unique(sym)) >= 4
    want = cpu.reconstruct(sym)
    st = torch.from_numpy(sym).to(cuda)
    fp32 = _with_impl(1, lambda: dec.reconstruct(st).cpu().numpy())
    tc = _with_impl(0, lambda: dec.reconstruct(st).cpu().numpy())
    assert want.shape == fp32.shape == tc.shape == (1, 3, 256, 512)
    assert np.abs(fp32 - want).max() < 1e-3, np.abs(fp32 - want).max()
    err = np.abs(tc - want)
    psnr = 10 * np.log10(1.0 / max(float(np.mean((tc - want) ** 2)), 1e-20))
    assert err.max() < 5e-2 and psnr > 40.0, (err.max(), psnr)
    # the reconstruction is not degenerate: it depends on the symbols
    other = dec.reconstruct(torch.roll(st, 7, dims=3)).cpu().numpy()
    assert np.abs(other - tc).max() > 1e-3


def test_widths_that_do_not_double(both, cuda):
    """W = 1664: the code is 104 columns wide and the round-half-up band widths do not double from scale to scale (band 0: 24 ->
    49 -> ...).  The reference leaves the convolutions in front of Dtow untrimmed (model_zoo_v2.py:166-173), so the last valid
    up-sampled column comes from one column beyond the coarse band; both product paths must follow (the port does, literally)."""
    import torch
    enc, dec, cpu, _ = both
    Hs, Ws = 256, 1664
    assert [int(v) for v in cpu.wl(1, Ws // 16)][:2] == [24, 50] and [int(v) for v in cpu.wl(2, Ws // 8)][0] == 49
    x = smooth_images(1, 3, Hs, Ws, seed=6)
    sym = cpu.symbols(x)
    want = cpu.reconstruct(sym)
    st = torch.from_numpy(sym).to(cuda)
    fp32 = _with_impl(1, lambda: dec.reconstruct(st).cpu().numpy())
    tc = _with_impl(0, lambda: dec.reconstruct(st).cpu().numpy())
    assert np.abs(fp32 - want).max() < 1e-3, np.abs(fp32 - want).max()
    assert np.abs(tc - want).max() < 5e-2, np.abs(tc - want).max()
    lat = cpu.analysis(x)
    xt = torch.from_numpy(x).to(cuda)
    assert np.abs(_with_impl(1, lambda: enc.latent(xt).cpu().numpy()) - lat).max() < 2e-4
    assert np.abs(_with_impl(0, lambda: enc.latent(xt).cpu().numpy()) - lat).max() < 2e-2


@pytest.mark.parametrize("Hs,Ws", [(1024, 2048), (2048, 4096)])
def test_large_image_round_trip(both, cuda, tmp_path, Hs, Ws):
    """Sizes beyond the reference's hard-wired 512x1024 (pseudo_codec.py:206-209): one-shot encoder with automatic slabs,
    the persistent dataflow decoder over 396 / 780 steps, offsets beyond 2^24 - symbols must come back bit for bit, from both
    decoder engines, and the reconstruction must be finite."""
    import torch
    from pseudocylindrical_convolution_b200 import _lib
    enc, dec, _, _ = both
    x = torch.from_numpy(smooth_images(1, 3, Hs, Ws, seed=11)).to(cuda)
    path = str(tmp_path / "big.bin")
    sym = enc.symbols(x)
    assert tuple(sym.shape) == (16, VD // 4, Hs // 128, Ws // 8)
    enc.ent.encode_batch(sym.clone(), [path])
    want = enc.ent.fill(sym.clone())
    assert os.path.getsize(path) * 8 / (Hs * Ws) > 0.05
    lib = _lib.load()
    try:
        for eng in (2, 1):
            lib.pcx_wave_set_fused(eng)
            got = dec.ent.decode_batch(Hs // 128, Ws // 8, [path])
            assert torch.equal(got, want), "engine %d" % eng
    finally:
        lib.pcx_wave_set_fused(2)
    rec = dec.decode_batch([path], Hs, Ws)
    assert tuple(rec.shape) == (1, 3, Hs, Ws) and torch.isfinite(rec).all()
    del rec, x
    torch.cuda.empty_cache()
