"""Depth <-> space shuffle (reference: PCONV_operator/Dtow.py)."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule


class Dtow(BaseOpModule):

    def __init__(self, stride=2, d2w=False, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.DtowOp(stride, d2w, gid, time_it) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]
