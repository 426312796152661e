"""ERP image -> 16 latitude bands (reference: PCONV_operator/SphereSlice.py)."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule
from .base import set_weight


class SphereSlice(BaseOpModule):
    """(N,C,H,W) -> (N*npart, C, H/npart + 2*pad, W + 2*pad); columns >= band width are zero."""

    def __init__(self, npart, interp_type=0, pad=0, opt=False, device=0, time_it=False):
        super().__init__(device)
        weight = set_weight(npart, opt)
        self.op = {gid: PCONV.SphereSliceOp(npart, interp_type, pad, weight, gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]
