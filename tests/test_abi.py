"""CPU: the C-ABI library builds, loads, and exports exactly what include/b200q.h declares
(no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g

    g.build()
    from pennylane_b200 import _lib

    return _lib.load()


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b200q.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200q_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from pennylane_b200 import _lib

    declared = _header_functions()
    assert declared, "no functions parsed from include/b200q.h"
    assert sorted(_lib.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared but not exported"


def test_exports_are_plain_c_symbols(lib):
    from pennylane_b200 import _lib

    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.SIGNATURES:
        getattr(raw, name)           # raises AttributeError if mangled / missing


def test_trivial_calls_without_gpu(lib):
    assert lib.b200q_version() == 1
    assert lib.b200q_workspace_bytes() >= (16 << 20)
    assert isinstance(lib.b200q_last_error(), bytes)


def test_argument_validation_fails_loudly(lib):
    """Bad arguments are rejected on the host before any launch (rc != 0, message set)."""
    from pennylane_b200._lib import B200QError, check, int_array

    tgt = int_array([5])
    rc = lib.b200q_apply_matrix(None, 3, 1, 1, tgt, 1, None, None, 0, None, None, 0, None)
    assert rc != 0
    with pytest.raises(B200QError, match="no matrix|out of range"):
        check(rc)
    rc = lib.b200q_apply_matrix(None, 3, 7, 1, int_array([0]), 1, None, None, 0,
                                ctypes.c_void_p(1), None, 0, None)
    assert rc != 0 and b"dtype" in lib.b200q_last_error()


def test_statevector_requires_cuda():
    import torch

    from pennylane_b200 import B200QError, StateVector

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(B200QError, match="CUDA"):
        StateVector(3)
