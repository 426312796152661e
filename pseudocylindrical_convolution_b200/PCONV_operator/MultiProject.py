"""14 rectilinear viewports of an ERP image (reference: PCONV_operator/MultiProject.py) - the sampling on which
`pseudo_codec.py --test` measures PSNR / SSIM (pseudo_codec.py:270-284).  Forward only."""
from .. import PCONV
from ._common import contiguous
from .BaseOpModule import BaseOpModule


class MultiProjectM(BaseOpModule):
    """viewports at caller-supplied angles (units of pi)"""

    def __init__(self, h, w, thetas, phis, fov=0.6, near=False, device_id=0, time_flag=False):
        super().__init__(device_id)
        self.op = {gid: PCONV.ProjectsOp(int(h), int(w), thetas, phis, fov, near, gid, time_flag) for gid in self.device_list}

    def forward(self, x):
        return self.native(x).forward(contiguous(x))[0]


class MultiProject(MultiProjectM):
    """the reference's fixed set: 4 x 3 viewports around the equator belt (+-45 degrees) and the two poles (MultiProject.py:39-40)"""

    def __init__(self, h, w, fov=0.6, near=False, device_id=0, time_flag=False):
        self.thetas = [-0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, -0.5, 0, 0.5, 1, 0, 0]
        self.phis = [0, 0, 0, 0, 0.25, 0.25, 0.25, 0.25, -0.25, -0.25, -0.25, -0.25, 0.5, -0.5]
        super().__init__(h, w, self.thetas, self.phis, fov, near, device_id, time_flag)
