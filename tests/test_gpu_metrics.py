"""`pseudo_codec.py --test` evaluation path: the 14-viewport projector against the UNMODIFIED reference
extension on the same GPU, the SSIM / MSE kernels against a plain PyTorch fp32 evaluation of the reference's formula
(PCONV_operator/pytorch_ssim.py:17-37), and the command line end to end (--enc, --dec, --test) on synthetic images with
random-init checkpoints.  Metric code: tolerances are stated, nothing here has to be bit-exact."""
import os

import numpy as np
import pytest

from conftest import smooth_images

pytestmark = pytest.mark.gpu

THETAS = [-0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, 0, 0]
PHIS = [0, 0, 0, 0, 0.25, 0.25, 0.25, 0.25, -0.25, -0.25, -0.25, -0.25, 0.5, -0.5]


@pytest.mark.parametrize("near", [False, True])
def test_viewport_projector_vs_reference(ref_ext, cuda, near):
    import torch
    from pseudocylindrical_convolution_b200 import PCONV
    if ref_ext is None:
        pytest.skip("oracle/_ref/PCONV_ref.so not available")
    x = torch.from_numpy(smooth_images(2, 3, 512, 1024, seed=8)).to(cuda)
    mine = PCONV.ProjectsOp(171, 256, THETAS, PHIS, 0.5, near, 0, False).forward(x)[0]
    ref = ref_ext.ProjectsOp(171, 256, THETAS, PHIS, 0.5, near, 0, False).forward(x)[0]
    assert tuple(mine.shape) == tuple(ref.shape) == (28, 3, 171, 256)
    diff = (mine - ref).abs()
    if near:
        # nearest sampling: a coordinate that differs in the last ulp can pick the neighbouring pixel at a .5 boundary
        assert float((diff > 0).float().mean()) < 1e-4
    else:
        assert float(diff.max()) < 2e-4, float(diff.max())          # sample coordinates agree to ~1e-4 pixel
    assert float(mine.min()) >= 0.0 and float(mine.max()) <= 1.0


def _torch_ssim(a, b, window=11, sigma=1.5):
    import math
    import torch
    import torch.nn.functional as F
    g = torch.tensor([math.exp(-(x - window // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window)])
    g = (g / g.sum()).unsqueeze(1)
    c = a.shape[1]
    w = g.mm(g.t()).float()[None, None].expand(c, 1, window, window).contiguous().to(a.device)
    conv = lambda t: F.conv2d(t, w, padding=window // 2, groups=c)
    mu1, mu2 = conv(a), conv(b)
    s1, s2, s12 = conv(a * a) - mu1 * mu1, conv(b * b) - mu2 * mu2, conv(a * b) - mu1 * mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))).mean()


def test_ssim_and_mse_kernels_vs_pytorch(cuda):
    import torch
    from pseudocylindrical_convolution_b200.PCONV_operator import SSIM
    from pseudocylindrical_convolution_b200.PCONV_operator.pytorch_ssim import mean_squared_difference
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        for shape, noise in (((14, 3, 171, 256), 0.05), ((2, 1, 37, 53), 0.3), ((1, 3, 11, 11), 0.0)):
            a = torch.from_numpy(smooth_images(*shape, seed=3)).to(cuda)
            b = (a + noise * torch.randn(a.shape, device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))).clamp(0, 1)
            got = float(SSIM(11, shape[1])(a, b))
            want = float(_torch_ssim(a, b))
            assert abs(got - want) < 2e-5, (shape, got, want)
            if noise == 0.0:
                assert abs(got - 1.0) < 1e-6
            mse = float(mean_squared_difference(a, b))
            assert abs(mse - float(((a - b) ** 2).double().mean())) < 1e-9 + 1e-6 * mse
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_command_line_enc_dec_test(cuda, tmp_path, monkeypatch, capsys):
    """README.md:13-22 of the reference: --enc writes one code file per image, --dec one image per code file, --test prints
    bitrate / PSNR / SSIM.  Runs in a scratch directory with synthesized ./demo/{mse,ssim} checkpoints."""
    import cv2
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    monkeypatch.chdir(tmp_path)
    synthesize_checkpoints("./demo/ssim", "4_56", 56, 0, seed=0)
    synthesize_checkpoints("./demo/ssim", "1_56", 56, 0, seed=0)       # check_models() looks for index 0 of both lists
    synthesize_checkpoints("./demo/mse", "1_56", 56, 0, seed=0)
    imgs = []
    for i in range(2):
        arr = (smooth_images(1, 3, 512, 1024, seed=40 + i)[0].transpose(1, 2, 0) * 255).astype(np.uint8)
        fn = str(tmp_path / ("img%d.png" % i))
        cv2.imwrite(fn, arr)
        imgs.append(fn)
    codes = [str(tmp_path / ("code%d.bin" % i)) for i in range(2)]
    outs = [str(tmp_path / ("rec%d.png" % i)) for i in range(2)]
    pc.main(["--enc", "--ssim", "--model-idx", "3", "--img-list", *imgs, "--code-list", *codes])
    assert all(os.path.getsize(c) > 1000 for c in codes)
    pc.main(["--dec", "--ssim", "--model-idx", "3", "--code-list", *codes, "--out-list", *outs])
    for o in outs:
        rec = cv2.imread(o)
        assert rec is not None and rec.shape == (512, 1024, 3)
    capsys.readouterr()
    pc.main(["--test", "--ssim", "--model-idx", "3", "--code-list", *codes, "--img-list", *imgs])
    text = capsys.readouterr().out
    assert "Average Performance" in text and "PSNR:" in text and "SSIM:" in text
    last = text.strip().splitlines()[-1]                                 # Bitrate:x.xxxbpp, PSNR:xx.xxdB, SSIM:x.xxxx
    rate = float(last.split("Bitrate:")[1].split("bpp")[0])
    psnr = float(last.split("PSNR:")[1].split("dB")[0])
    ssim = float(last.split("SSIM:")[1])
    assert abs(rate - np.mean([os.path.getsize(c) * 8 / 1024. / 512. for c in codes])) < 1e-3
    assert 3.0 < psnr < 60.0 and -1.0 <= ssim <= 1.0                     # random-init weights: only sanity, not quality


def test_command_line_other_sizes(cuda, tmp_path, monkeypatch, capsys):
    """--height / --width: the reference resizes everything to 1024 x 512 and hard-wires the code size
    (pseudo_codec.py:206-209, :229-234); here the coded size is an option on all three commands, default unchanged.  Inputs of
    another size are resized to it, the bitrate is reported per coded pixel, a bad size is refused."""
    import cv2
    from pseudocylindrical_convolution_b200 import pseudo_codec as pc
    from pseudocylindrical_convolution_b200.random_init import synthesize_checkpoints
    monkeypatch.chdir(tmp_path)
    for d, p in (("./demo/ssim", "4_56"), ("./demo/ssim", "1_56"), ("./demo/mse", "1_56")):
        synthesize_checkpoints(d, p, 56, 0, seed=0)
    arr = (smooth_images(1, 3, 300, 700, seed=9)[0].transpose(1, 2, 0) * 255).astype(np.uint8)      # not the coded size: resized
    img, code, out = str(tmp_path / "a.png"), str(tmp_path / "a.bin"), str(tmp_path / "a_rec.png")
    cv2.imwrite(img, arr)
    size = ["--height", "256", "--width", "768"]
    pc.main(["--enc", "--ssim", "--model-idx", "3", "--img-list", img, "--code-list", code] + size)
    text = capsys.readouterr().out
    rate = float(text.split("bitrate:")[1].split("bpp")[0])
    assert abs(rate - os.path.getsize(code) * 8 / (256 * 768)) < 1e-3
    pc.main(["--dec", "--ssim", "--model-idx", "3", "--code-list", code, "--out-list", out] + size)
    rec = cv2.imread(out)
    assert rec is not None and rec.shape == (256, 768, 3)
    pc.main(["--test", "--ssim", "--model-idx", "3", "--code-list", code, "--img-list", img] + size)
    assert "Average Performance" in capsys.readouterr().out
    with pytest.raises(AssertionError):
        pc.main(["--enc", "--ssim", "--model-idx", "3", "--img-list", img, "--code-list", code, "--height", "300", "--width", "768"])
