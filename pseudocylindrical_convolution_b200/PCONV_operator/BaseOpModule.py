"""Device plumbing shared by the operator modules (reference: PCONV_operator/BaseOpModule.py:5-54).

A module keeps one native op object per GPU id in `self.op`; moving the module re-keys the entry and tells the
native object (`op.to(id)`), exactly like the reference.  nn.DataParallel replication shares the op table.
"""
from collections import OrderedDict

from torch import nn


class BaseOpModule(nn.Module):

    def __init__(self, devices=0):
        super().__init__()
        self.device_list = [devices] if isinstance(devices, int) else list(devices)
        self.apply_flag = False

    # nn.Module.to() funnels through _apply(fn); remember that it ran so the custom to() can re-key the op
    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        self.apply_flag = True
        fn(self)
        return self

    def custom_op_replicate(self, other):
        other.op = self.op
        return other

    def _replicate_for_data_parallel(self):
        replica = self.custom_op_replicate(self.__new__(type(self)))
        replica.__dict__ = self.__dict__.copy()
        replica._parameters = OrderedDict()
        replica._buffers = replica._buffers.copy()
        replica._modules = replica._modules.copy()
        replica._is_replica = True
        return replica

    def custom_op_to(self, *args):
        if args and args[0] is not None and getattr(args[0], "index", None) is not None and len(self.op) == 1:
            new_id, old_id = args[0].index, next(iter(self.op))
            if new_id != old_id:
                self.op[new_id] = self.op.pop(old_id)
                self.op[new_id].to(new_id)

    # fn(self) above calls these on the module itself (torch probes tensors this way)
    def is_floating_point(self):
        return False

    def is_complex(self):
        return False

    def to(self, *args, **kwargs):
        if not self.apply_flag:
            super().to(*args, **kwargs)
        else:
            self.custom_op_to(*args)
            self.apply_flag = False
        return self

    # helper used by every wrapper: native op bound to the tensor's device
    def native(self, x):
        try:
            return self.op[x.device.index]
        except KeyError:
            raise RuntimeError("%s has no native op for %s (built for GPUs %s)" % (type(self).__name__, x.device, list(self.op)))
