"""Device bookkeeping shared by the operator modules.

Contract kept from the reference (PCONV_operator/BaseOpModule.py:5-54): a module owns one native op object per GPU id in the
dict `self.op`; moving the module to another GPU (`.to('cuda:1')`, `.cuda(1)`) re-keys its single entry and tells the native
object (`op.to(id)`); `device_list` records the ids it was built for; replicas made by nn.DataParallel share the op table.

How it is done here: nn.Module funnels every move through `_apply(fn)`.  Instead of posing as a tensor, the module asks `fn` where
it sends things by passing it an empty probe tensor, and follows when the answer is a CUDA device.
"""
import torch
from torch import nn


class BaseOpModule(nn.Module):

    def __init__(self, devices=0):
        super().__init__()
        self.device_list = [devices] if isinstance(devices, int) else [int(d) for d in devices]

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        target = self._target_gpu(fn)
        if target is not None:
            self._follow(target)
        return self

    @staticmethod
    def _target_gpu(fn):
        """GPU index `fn` moves tensors to, or None (dtype casts, .cpu(), share_memory_() ...)"""
        try:
            dev = fn(torch.empty(0)).device
        except Exception:
            return None
        return dev.index if dev.type == "cuda" and dev.index is not None else None

    def _follow(self, gpu):
        ops = getattr(self, "op", None)
        if not isinstance(ops, dict) or len(ops) != 1:      # built for several GPUs: every id keeps its own op
            return
        (old, op), = ops.items()
        if old != gpu:
            del ops[old]
            ops[gpu] = op
            op.to(gpu)
            self.device_list = [gpu]

    def native(self, x):
        """the native op bound to the device of tensor x"""
        op = self.op.get(x.device.index)
        if op is None:
            raise RuntimeError("%s has no native op for %s (built for GPUs %s)" % (type(self).__name__, x.device, sorted(self.op)))
        return op
