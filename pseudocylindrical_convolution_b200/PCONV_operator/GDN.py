"""LowerBound (reference: PCONV_operator/GDN.py:6-22) - forward only."""
import torch


class LowerBound:
    @staticmethod
    def apply(inputs, bound):
        return torch.clamp(inputs, min=float(bound))
