"""Gaussian-window SSIM (reference: PCONV_operator/pytorch_ssim.py), evaluated by one CUDA kernel (pcx_ssim) instead of five
grouped F.conv2d calls.  Also the mean-squared-difference reduction `--test` turns into PSNR."""
import ctypes as C

import torch

from .._lib import call


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _reduce_scratch(total, device):
    blocks = (total + 255) // 256
    return torch.empty((blocks + 1,), dtype=torch.float64, device=device), blocks


def ssim(img1, img2, window_size=11, size_average=True):
    if not size_average:
        raise NotImplementedError("per-image SSIM (size_average=False) is not used by the codec's --test path")
    if img1.shape != img2.shape or img1.dim() != 4:
        raise ValueError("ssim expects two (N, C, H, W) tensors of the same shape")
    a, b = img1.contiguous().float(), img2.contiguous().float()
    n, c, h, w = a.shape
    with torch.cuda.device(a.device):
        scratch, blocks = _reduce_scratch(a.numel(), a.device)
        call("pcx_ssim", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), n * c, h, w, int(window_size), C.c_float(1.5), None,
             C.c_void_p(scratch.data_ptr()), blocks, C.c_void_p(scratch.data_ptr() + 8 * blocks), _stream())
        return scratch[blocks].to(torch.float32)


class SSIM(torch.nn.Module):
    def __init__(self, window_size=11, channel=1, size_average=True):
        super().__init__()
        self.window_size, self.channel, self.size_average = window_size, channel, size_average

    def forward(self, img1, img2):
        return ssim(img1, img2, self.window_size, self.size_average)


def mean_squared_difference(x, y):
    """torch.mean((x - y) ** 2) of pseudo_codec.py:276 as one deterministic two-pass reduction"""
    a, b = x.contiguous().float(), y.contiguous().float()
    with torch.cuda.device(a.device):
        scratch, blocks = _reduce_scratch(a.numel(), a.device)
        call("pcx_mean_sqdiff", C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), a.numel(), C.c_void_p(scratch.data_ptr()), blocks,
             C.c_void_p(scratch.data_ptr() + 8 * blocks), _stream())
        return scratch[blocks].to(torch.float32)
