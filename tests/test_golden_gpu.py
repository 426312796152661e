"""GPU: the product (PCONV mirror -> C ABI -> libpcx.so) against the committed golden vectors, which are outputs of
the UNMODIFIED reference extension on a B200 (tests/golden/PROVENANCE.txt).  Bit-exact everywhere: these are
CUDA-core fp32 / integer paths with pinned expression shapes, and the CDF tables feed the arithmetic coder."""
import os

import numpy as np
import pytest

import golden_cases as gc
from test_gpu_parity import assert_bit_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    path = os.path.join(GOLD, "ref_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden file %s missing" % path)
    return np.load(path)


@pytest.mark.parametrize("name", ["tiles", "quant", "gmm"])
def test_product_matches_reference_outputs(cuda, name):
    from pseudocylindrical_convolution_b200 import PCONV
    want = _gold(name)
    got = gc.CASES[name](PCONV, cuda)
    assert set(got) == set(want.files)
    for k in want.files:
        assert_bit_equal(got[k], want[k], "%s/%s" % (name, k))


def test_product_wavefront_matches_reference_stream(cuda, tmp_path):
    """CDF rows of every wavefront step, the extracted GMM parameters (digest) and the bitstream bytes."""
    from pseudocylindrical_convolution_b200 import PCONV, coder
    want = _gold("wavefront")
    got = gc.case_wavefront(PCONV, cuda, coder, tmp_path / "mine.bin")
    for k in ("counts", "labels", "digests", "tables", "bitstream"):
        assert_bit_equal(got[k], want[k], "wavefront/%s" % k)
