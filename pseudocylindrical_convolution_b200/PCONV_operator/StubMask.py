"""Channel masks / channel slice (reference: PCONV_operator/StubMask.py)."""
import torch
from torch import nn


class StubMask(nn.Module):
    """Returns a {0,1} mask that keeps the first `dims` channels."""

    def __init__(self, dims=192):
        super().__init__()
        self.dims = dims
        self.mask = None

    def forward(self, x):
        if self.mask is None or self.mask.shape != x.shape or self.mask.device != x.device:
            self.mask = torch.ones_like(x)
            self.mask[:, self.dims:] = 0
        return self.mask


class Extract(nn.Module):
    """x[:, :dims] as a contiguous tensor."""

    def __init__(self, dims):
        super().__init__()
        self.dims = dims

    def forward(self, x):
        return x[:, :self.dims].contiguous()
