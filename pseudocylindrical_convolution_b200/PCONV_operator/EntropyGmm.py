"""GMM negative log-likelihood of the symbols (reference: PCONV_operator/EntropyGmm.py) - rate estimate, forward only."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule


class EntropyGmm(BaseOpModule):

    def __init__(self, num_gaussian=3, ignore_label=0, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.EntropyGmmOp(num_gaussian, ignore_label, gid, time_it) for gid in self.device_list}

    def forward(self, weight, delta, mean, label):
        return self.native(weight).forward(contiguous(weight), contiguous(delta), contiguous(mean), contiguous(label))[0]
