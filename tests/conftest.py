import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box)")


def _has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def random_state(n, seed=0, batch=None, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    shape = (2,) * n if batch is None else (batch,) + (2,) * n
    st = rng.normal(size=shape) + 1j * rng.normal(size=shape)
    if batch is None:
        st /= np.linalg.norm(st)
    else:
        st /= np.linalg.norm(st.reshape(batch, -1), axis=1).reshape((batch,) + (1,) * n)
    return st.astype(dtype)


TOL = {np.dtype(np.complex128): 1e-12, np.dtype(np.complex64): 1e-5}
