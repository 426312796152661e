"""`PCONV_operator` - nn.Module wrappers, one per native operator (reference: PCONV_operator/__init__.py).

Same public names as the reference package for everything on the codec hot path and for the `--test` metrics
(MultiProject, SSIM).  Training-only utilities (MaskConv2, ContextReshape, DropGrad, ModuleSaver, Logger) are out of
scope; see DESIGN.md.
"""
from .Dtow import Dtow
from .EntropyGmm import EntropyGmm
from .SphereSlice import SphereSlice
from .SphereUslice import SphereUslice
from .StubMask import StubMask, Extract
from .EntropyGmmTable import EntropyGmmTable, EntropyBatchGmmTable
from .EntropyContextNew import (EntropyContextNew, EntropyConv2, EntropyConv2Batch, EntropyCtxPadRun2, DExtract2, DInput2,
                                DExtract2Batch, EntropyAdd)
from .PseudoContextV2 import (PseudoFillV2, PseudoContextV2, PseudoGDNV2, PseudoPadV2, PseudoEntropyContext,
                              PseudoEntropyPad, PseudoQUANTV2, PseudoDQUANT)
from .base import set_weight
from .MultiProject import MultiProject, MultiProjectM
from .pytorch_ssim import SSIM
