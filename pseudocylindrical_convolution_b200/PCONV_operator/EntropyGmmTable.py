"""GMM parameters -> integer CDF tables for the arithmetic coder (reference: PCONV_operator/EntropyGmmTable.py)."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule


class EntropyGmmTable(BaseOpModule):

    def __init__(self, nstep, bias, num_gaussian, total_region, beta=1e-6, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.EntropyGmmTableOp(nstep, bias, num_gaussian, total_region, beta, gid, time_it) for gid in self.device_list}

    def forward(self, weight, delta, mean, ntop):
        return self.native(weight).forward(contiguous(weight), contiguous(delta), contiguous(mean), ntop)[0]


class EntropyBatchGmmTable(BaseOpModule):
    """Input = cat([mixture logits, delta, mean]) as produced by DExtract2Batch; output (rows, nstep+1)."""

    def __init__(self, nstep, bias, num_gaussian, total_region, beta=1e-6, device=0, time_it=False):
        super().__init__(device)
        self.op = {gid: PCONV.EntropyGmmTableOp(nstep, bias, num_gaussian, total_region, beta, gid, time_it) for gid in self.device_list}

    def forward(self, x, ntop):
        return self.native(x).forward_batch(contiguous(x), ntop)[0]
