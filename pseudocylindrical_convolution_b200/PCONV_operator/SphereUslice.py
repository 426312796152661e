"""Latitude bands -> ERP image (reference: PCONV_operator/SphereUslice.py)."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule
from .base import set_weight


class SphereUslice(BaseOpModule):

    def __init__(self, npart, interp_type=0, pad=0, opt=False, device=0, time_it=False):
        super().__init__(device)
        weight = set_weight(npart, opt)
        self.op = {gid: PCONV.SphereUsliceOp(npart, interp_type, pad, weight, gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]
