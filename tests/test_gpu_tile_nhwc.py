"""GPU parity of the fused channels-last tile gathers and of the fused tile pipeline (BASELINE.json config 2).

slice+pad and uslice only move and interpolate values with the reference's expression shapes, so they are
compared BIT-EXACTLY with the CPU oracle's slice -> pad / uslice on the same inputs; the fused pipeline contains
the TF32 tensor-core convolution and is compared with the fp32 chain under a stated tolerance."""
import ctypes as C

import numpy as np
import pytest

from conftest import W64, smooth_images
from test_gpu_parity import N, T, assert_bit_equal

pytestmark = pytest.mark.gpu
WEIGHT = [float(v) for v in W64]


def _tables(P, cuda, H, W, pad):
    import torch
    sl = P.SphereSliceOp(16, 0, 0, WEIGHT, 0, False)
    us = P.SphereUsliceOp(16, 0, 0, WEIGHT, 0, False)
    like = torch.zeros(1, device=cuda)
    wl, s_src, s_wt = sl._geometry(H, W, like, "pcx_slice_table")
    _, u_src, u_wt = us._geometry(H, W, like, "pcx_uslice_table")
    halo = (None, None, None, None)
    ctx = P.PseudoContextOp(16, 20, WEIGHT, 0, False)
    if pad > 0:
        halo = ctx.halo(1, H // 16, W, pad)
    return wl, (s_src, s_wt), (u_src, u_wt), halo, (sl, us, ctx)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


@pytest.mark.parametrize("n,c,H,W,pad,extra", [
    (1, 3, 64, 128, 1, 0),        # whole rows staged (W <= 320), 3 channels
    (2, 40, 32, 64, 2, 3),        # two images, ragged channel chunk, pad 2, wider pitch
    (1, 32, 64, 1024, 1, 0),      # chunked columns, polar bands span ~4.3 source columns per column
    (1, 33, 32, 2048, 2, 1),
    (1, 8, 128, 512, 0, 0),       # no halo
    (1, 4, 16, 4096, 1, 0),       # one row per band: every halo row comes from another band or the pole mirror
])
def test_slice_pad_nhwc_bit_exact(P, cuda, orc, n, c, H, W, pad, extra):
    import torch
    from pseudocylindrical_convolution_b200._lib import call, int_array
    x = smooth_images(n, c, H, W, seed=5)
    wl, (s_src, s_wt), _, (band, row, col, tw), keep = _tables(P, cuda, H, W, pad)
    assert wl == list(orc.band_widths(W64, H, W))
    h = H // 16
    pitch = W + 2 * pad + extra
    out = torch.full((n * 16, h + 2 * pad, pitch, c), 123.0, device=cuda)
    call("pcx_slice_pad_nhwc", _ptr(T(x, cuda)), _ptr(out), n, c, H, W, 16, pad, int_array(wl), _ptr(s_src), _ptr(s_wt),
         _ptr(band), _ptr(row), _ptr(col), _ptr(tw), pitch, 1, None)
    torch.cuda.synchronize()
    want = orc.sphere_slice(x, wl)
    if pad > 0:
        want = orc.pseudo_pad(want, wl, pad)
    got = N(out).transpose(0, 3, 1, 2)
    assert_bit_equal(got[:, :, :, :W + 2 * pad], want, "slice+pad nhwc")
    assert (got[:, :, :, W + 2 * pad:] == 0).all()
    # zero_invalid = 0 leaves the columns beyond the bands untouched
    out2 = torch.full((n * 16, h + 2 * pad, pitch, c), 123.0, device=cuda)
    call("pcx_slice_pad_nhwc", _ptr(T(x, cuda)), _ptr(out2), n, c, H, W, 16, pad, int_array(wl), _ptr(s_src), _ptr(s_wt),
         _ptr(band), _ptr(row), _ptr(col), _ptr(tw), pitch, 0, None)
    got2 = N(out2).transpose(0, 3, 1, 2)
    for g in range(16):
        assert_bit_equal(got2[g::16, :, :, :wl[g] + 2 * pad], want[g::16, :, :, :wl[g] + 2 * pad], "band %d" % g)
        assert (got2[g::16, :, :, wl[g] + 2 * pad:] == 123.0).all()


@pytest.mark.parametrize("n,c,h,W,y0,x0,extra", [
    (1, 3, 4, 128, 0, 0, 0),
    (2, 40, 2, 64, 1, 2, 1),
    (1, 32, 4, 1024, 0, 0, 0),
    (1, 33, 2, 2048, 2, 2, 3),
    (1, 4, 1, 4096, 1, 1, 0),
])
def test_uslice_nhwc_bit_exact(P, cuda, orc, n, c, h, W, y0, x0, extra):
    import torch
    from pseudocylindrical_convolution_b200._lib import call, int_array
    rng = np.random.default_rng(11)
    wl, _, (u_src, u_wt), _, keep = _tables(P, cuda, 16 * h, W, 0)
    tiles = orc.pseudo_fill(rng.standard_normal((n * 16, c, h, W)).astype(np.float32), wl)
    rows, pitch = h + 2 * y0, W + 2 * x0 + extra
    big = rng.standard_normal((n * 16, rows, pitch, c)).astype(np.float32)       # garbage around the window
    big[:, y0:y0 + h, x0:x0 + W, :] = tiles.transpose(0, 2, 3, 1)
    out = torch.full((n, c, 16 * h, W), -5.0, device=cuda)
    call("pcx_uslice_nhwc", _ptr(T(big, cuda)), _ptr(out), n, c, h, W, 16, rows, pitch, y0, x0, int_array(wl),
         _ptr(u_src), _ptr(u_wt), None)
    torch.cuda.synchronize()
    assert_bit_equal(N(out), orc.sphere_uslice(tiles, wl), "uslice nhwc")


@pytest.mark.parametrize("Ci,Co,H,W,act", [(32, 96, 64, 128, False), (192, 192, 128, 256, True), (64, 192, 32, 1024, True)])
def test_tile_pipeline_vs_fp32_chain(cuda, orc, Ci, Co, H, W, act):
    """Fused slice+pad -> tcgen05 conv (+PReLU, fill) -> uslice against the same chain through the separate NCHW
    operators with the fp32 CUDA-core convolution.  The gathers are exact; the only difference is TF32 operand
    precision in the convolution (10 mantissa bits): rms error < 1.5e-3, max < 1e-2 for unit-variance outputs."""
    import torch
    from pseudocylindrical_convolution_b200.tile_pipeline import TilePipeline, reference_chain
    torch.manual_seed(3)
    pipe = TilePipeline(Ci, Co, act=act, device=0)
    with torch.no_grad():
        pipe.conv.weight.normal_(0, 1.0 / np.sqrt(9 * Ci))
        pipe.conv.bias.normal_(0, 0.1)
        if act:
            pipe.relu.weight.uniform_(0.05, 0.5)
    x = T(smooth_images(2, Ci, H, W, seed=9) * 2 - 1, cuda)
    got = pipe(x)
    want = reference_chain(x, pipe.conv, pipe.relu, device=0, impl=1)
    torch.cuda.synchronize()
    err = (got - want).abs()
    assert float(err.pow(2).mean().sqrt()) < 1.5e-3
    assert float(err.max()) < 1e-2
    # same call on pinned host buffers (streams + double buffering) gives the same bits as the device call
    xh = [x[i].cpu().pin_memory() for i in range(2)]
    oh = [torch.empty((Co, H, W)).pin_memory() for _ in range(2)]
    pipe.forward_host(xh, oh)
    torch.cuda.synchronize()
    for i in range(2):
        assert torch.equal(oh[i], got[i].cpu())


@pytest.fixture(scope="module")
def P(cuda):
    from pseudocylindrical_convolution_b200 import PCONV
    return PCONV
