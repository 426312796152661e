import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W64 = [15, 31, 54, 63, 63, 64, 64, 64, 64, 64, 64, 63, 63, 54, 31, 15]   # set_weight(16, opt=True)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (oracle/oracle.py) - the checker, never the thing under test."""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")


def _load_ext(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        return None
    import torch  # noqa: F401  (loads libtorch / libc10 the extension links against)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def ref_ext():
    """The UNMODIFIED reference CUDA extension built for sm_100 (oracle/build_ref.py), or None."""
    try:
        return _load_ext("PCONV_ref")
    except Exception as e:                      # pragma: no cover
        print("reference extension not loadable:", e)
        return None


@pytest.fixture(scope="session")
def ref_coder():
    try:
        return _load_ext("coder_ref")
    except Exception as e:                      # pragma: no cover
        print("reference coder not loadable:", e)
        return None


def smooth_images(n, c, h, w, seed=1234):
    """Synthetic ERP content: smooth low-frequency field + noise in [0,1]."""
    rng = np.random.default_rng(seed)
    low = rng.random((n, c, max(h // 8, 1), max(w // 8, 1))).astype(np.float32)
    up = np.repeat(np.repeat(low, -(-h // low.shape[2]), axis=2), -(-w // low.shape[3]), axis=3)[:, :, :h, :w]
    return (0.8 * up + 0.2 * rng.random((n, c, h, w)).astype(np.float32)).astype(np.float32)
