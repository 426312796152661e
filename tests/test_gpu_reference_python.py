"""The reference's own Python layer on the GPU box, against this repository (route-A drop-in proof, bpp / PSNR
/ SSIM equality with the reference pipeline on the same random-init checkpoints).

oracle/build_ref.py stages the reference's Python modules as marshalled code objects in oracle/_ref/refpy and compiles its CUDA
extension + coder unmodified (PCONV_ref.so, coder_ref.so); tests/ref_runner.py runs the reference's PseudoEncoder /
PseudoDecoder in a fresh interpreter over either back end.  Skipped when those artefacts are absent.

  * `ref`    back end = the REAL reference end to end.  Its bitstream must decode with the product decoder to exactly the symbols
    it coded (every one of the 93,632 CDF rows has to agree bit for bit, or the range decoder derails), the product's bitstream
    must decode with the reference decoder, bpp agree up to the counted symbol flips of the TF32 transforms, reconstructions
    agree within the stated tolerance and viewport PSNR / SSIM agree.
  * `mirror` back end = the reference's Python over this repository's `PCONV` / `coder` modules (INTEGRATION.md route A): it must
    produce the reference back end's bitstream BYTE FOR BYTE - same cuDNN convolutions, bit-exact custom operators."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, smooth_images

pytestmark = pytest.mark.gpu
H, W, VD, PREX = 512, 1024, 56, "4_56"
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _have_ref():
    return all(os.path.exists(os.path.join(REFDIR, n)) for n in ("PCONV_ref.so", "coder_ref.so", "refpy/pseudo_codec.pcb"))


def _run(backend, models, image, out, decode=()):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_runner.py"), "--backend", backend, "--models", models, "--image", image, "--out", out]
    for name, path in decode:
        cmd += ["--decode", "%s=%s" % (name, path)]
    os.makedirs(out, exist_ok=True)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=out)
    assert r.returncode == 0, "ref_runner(%s) failed:\n%s\n%s" % (backend, r.stdout[-2000:], r.stderr[-4000:])


@pytest.fixture(scope="module")
def world(cuda, tmp_path_factory):
    """product codec + checkpoints (made by the CPU port's seeded generator, strict-loaded by both sides) + one image"""
    if not _have_ref():
        pytest.skip("oracle/_ref (reference extensions + byte-compiled Python layer) not built")
    import torch
    from oracle import cpu_codec as cc
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    d = str(tmp_path_factory.mktemp("refpy"))
    (p_enc, p_dec, p_ent), _ = cc.save_checkpoints(d, PREX, VD, seed=0)
    enc = pc.PseudoEncoder(VD, 0).to(cuda)
    dec = pc.PseudoDecoder(VD, 0).to(cuda)
    pc.load_models(enc, p_enc, p_ent, "cuda:0")                  # strict: the port's key set is the reference's
    pc.load_models(dec, p_dec, p_ent, "cuda:0")
    x = smooth_images(1, 3, H, W, seed=77)
    img = os.path.join(d, "img.npy")
    np.save(img, x)
    prod_bin = os.path.join(d, "prod.bin")
    xt = torch.from_numpy(x).to(cuda)
    enc(xt, prod_bin)
    return dict(dir=d, enc=enc, dec=dec, x=xt, img=img, prod_bin=prod_bin)


def _psnr(a, b):
    return 10 * np.log10(1.0 / max(float(np.mean((a - b) ** 2)), 1e-20))


def test_real_reference_and_product_decode_each_other(world):
    import torch
    from pseudocylindrical_convolution_b200.PCONV_operator import MultiProject, SSIM
    from pseudocylindrical_convolution_b200.PCONV_operator.pytorch_ssim import mean_squared_difference
    d, enc, dec, x = world["dir"], world["enc"], world["dec"], world["x"]
    out = os.path.join(d, "ref")
    _run("ref", d, world["img"], out, decode=[("own", os.path.join(out, "enc.bin")), ("prod", world["prod_bin"])])
    ref_sym = np.load(os.path.join(out, "sym.npy"))
    # (1) the reference decoder recovers what the reference encoder coded, and so does the PRODUCT decoder from the same bytes
    assert np.array_equal(np.load(os.path.join(out, "own_sym.npy")), ref_sym)
    dec.ent.start(os.path.join(out, "enc.bin"))
    got = dec.ent(H // 128, W // 8).cpu().numpy()
    assert np.array_equal(got, ref_sym), "product decoder derailed on the reference's bitstream"
    # (2) the reference decoder recovers the PRODUCT encoder's symbols from the product's bytes
    prod_sym = enc.ent.fill(enc.symbols(x).clone()).cpu().numpy()
    assert np.array_equal(np.load(os.path.join(out, "prod_sym.npy")), prod_sym), "reference decoder derailed on the product's bitstream"
    # (3) transforms: cuDNN (TF32 allowed by torch's default) vs tcgen05 TF32 - symbol flips near bin edges are counted
    flips = float((prod_sym != ref_sym).mean())
    lat_ref = np.load(os.path.join(out, "latent.npy"))
    lat = enc.latent(x).cpu().numpy()
    assert np.abs(lat - lat_ref).max() < 2e-2 and np.sqrt(np.mean((lat - lat_ref) ** 2)) < 2e-3
    assert flips < 0.02, flips
    n_ref, n_prod = os.path.getsize(os.path.join(out, "enc.bin")), os.path.getsize(world["prod_bin"])
    bpp_ref, bpp_prod = n_ref * 8 / (H * W), n_prod * 8 / (H * W)
    assert abs(bpp_ref - bpp_prod) <= 0.02 * bpp_ref + 3 * flips, (bpp_ref, bpp_prod, flips)
    # (4) same bytes through both decoders: reconstructions within tolerance, viewport PSNR / SSIM equal to 0.05 dB / 1e-3
    rec_ref = np.load(os.path.join(out, "prod_rec.npy"))
    rec = dec(world["prod_bin"], H, W)
    rn = rec.cpu().numpy()
    assert np.abs(rn - rec_ref).max() < 5e-2 and _psnr(rn, rec_ref) > 45.0, (_psnr(rn, rec_ref), np.abs(rn - rec_ref).max())
    pr = MultiProject(171, int(171 * 1.5), 0.5, False, 0).to(x.device)
    sim = SSIM(11, 3).to(x.device)
    vx = pr(x).clone()                                   # the op returns its cached output buffer (base_opt.hpp:43-57 semantics)
    res = {}
    for tag, r in (("prod", rec), ("ref", torch.from_numpy(rec_ref).to(x.device))):
        vy = pr(r.contiguous()).clone()
        res[tag] = (10 * np.log10(1.0 / mean_squared_difference(vx, vy).item()), sim(vx, vy).item())
    assert abs(res["prod"][0] - res["ref"][0]) < 0.05 and abs(res["prod"][1] - res["ref"][1]) < 1e-3, res
    print("bpp ref %.4f prod %.4f, symbol flips %.4f%%, viewport PSNR / SSIM prod %s ref %s" % (bpp_ref, bpp_prod, 100 * flips, res["prod"], res["ref"]))


def test_route_a_reference_python_over_the_mirror_is_byte_identical(world):
    """INTEGRATION.md route A: the unmodified reference Python (PCONV_operator/*.py, model_zoo_v2.py, pseudo_codec.py) with
    sys.modules['PCONV'] / ['coder'] bound to this repository's mirrors - versus the same Python over the reference's own
    extensions.  Both run cuDNN for the dense convolutions; every custom operator and the coder are the part under test."""
    d = world["dir"]
    out_ref, out_mir = os.path.join(d, "ref_a"), os.path.join(d, "mirror_a")
    _run("ref", d, world["img"], out_ref)
    _run("mirror", d, world["img"], out_mir, decode=[("own", os.path.join(out_ref, "enc.bin"))])
    a, b = np.load(os.path.join(out_ref, "latent.npy")), np.load(os.path.join(out_mir, "latent.npy"))
    assert np.array_equal(a, b), "analysis transform differs: max %g" % np.abs(a - b).max()
    assert np.array_equal(np.load(os.path.join(out_ref, "sym.npy")), np.load(os.path.join(out_mir, "sym.npy")))
    ra, rb = open(os.path.join(out_ref, "enc.bin"), "rb").read(), open(os.path.join(out_mir, "enc.bin"), "rb").read()
    assert len(ra) > 1000 and ra == rb, "bitstreams differ (%d vs %d bytes)" % (len(ra), len(rb))
    # and the reference's decoder loop over the mirror decodes the reference's stream to the reference's symbols
    assert np.array_equal(np.load(os.path.join(out_mir, "own_sym.npy")), np.load(os.path.join(out_ref, "sym.npy")))
