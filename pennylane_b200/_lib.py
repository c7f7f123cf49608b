"""ctypes binding of the C ABI declared in ``include/b200q.h``.

The shared library is built in-tree by ``__graft_entry__.build()`` (or ``python -m
pennylane_b200.build``) into ``pennylane_b200/csrc/libb200q.so``.  There is NO fallback: if the
library is missing or a call fails, an exception is raised — results never come from anywhere
but the CUDA kernels.
"""
from __future__ import annotations

import ctypes as C
import os
from functools import lru_cache

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libb200q.so")

_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_d = C.c_double
_p = C.c_void_p
_sz = C.c_size_t
_ip = C.POINTER(C.c_int)
_u64p = C.POINTER(C.c_uint64)
_dp = C.POINTER(C.c_double)

#: name -> (restype, argtypes).  This table IS the Python view of include/b200q.h; the
#: non-GPU test-suite checks that every symbol here is exported by the library and that every
#: function declared in the header is listed here.
SIGNATURES = {
    "b200q_last_error": (C.c_char_p, []),
    "b200q_version": (_i, []),
    "b200q_sm_count": (_i, []),
    "b200q_workspace_bytes": (_sz, []),
    "b200q_set_basis_state": (_i, [_p, _i, _i, _i64, _u64, _p]),
    "b200q_apply_matrix": (_i, [_p, _i, _i, _i64, _ip, _i, _ip, _ip, _i, _p, _p, _i64, _p]),
    "b200q_apply_diag": (_i, [_p, _i, _i, _i64, _ip, _i, _p, _p, _i64, _p]),
    "b200q_apply_phase": (_i, [_p, _i, _i, _i64, _ip, _ip, _i, _d, _d, _p, _p]),
    "b200q_gram_block": (_i, [_p, _i, _i, _ip, _i, _ip, _i, _u64, _u64, _p, _p, _sz, _p]),
    "b200q_collapse": (_i, [_p, _i, _i, _i, _i, _i, _d, _p]),
    "b200q_apply_parity_phase": (_i, [_p, _i, _i, _i64, _u64, _d, _d, _d, _d, _p, _p]),
    "b200q_apply_pauli_rot": (_i, [_p, _i, _i, _i64, _u64, _u64, _i, _d, _d, _p, _p]),
    "b200q_probs": (_i, [_p, _i, _i, _i64, _ip, _i, _p, _p, _sz, _p]),
    "b200q_expval_pauli_sum": (_i, [_p, _i, _i, _i64, _u64p, _u64p, _ip, _dp, _i, _p, _p, _sz, _p]),
    "b200q_inner": (_i, [_p, _p, _i, _i, _i64, _p, _p, _sz, _p]),
    "b200q_pauli_sum_apply": (
        _i, [_p, _p, _i, _i, _i64, _u64p, _u64p, _ip, _dp, _dp, _i, _d, _p, _sz, _p]),
    "b200q_pauli_braket": (_i, [_p, _p, _i, _i, _u64, _u64, _i, _p, _p, _sz, _p]),
    "b200q_sample": (_i, [_p, _i, _p, _i64, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "b200q_has_nan": (_i, [_p, _i, _p, _p]),
    "b200q_np_sum": (_i, [_p, _i, _p, _p, _sz, _p]),
    "b200q_div_by": (_i, [_p, _i, _p, _p]),
    "b200q_cumsum": (_i, [_p, _i, _i, _p, _p, _sz, _p]),
    "b200q_search": (_i, [_p, _i, _p, _i64, _p, _p]),
    "b200q_unpack_bits": (_i, [_p, _i64, _i, _p, _p]),
    "b200q_apply_tile": (_i, [_p, _i, _i, _i64, _ip, _i, _i, _p, _i, _p, _i, _p, _sz, _p]),
    "b200q_rtile_geometry": (_i, [_i, _i, _ip, _ip, _ip]),
    "b200q_apply_rtile": (_i, [_p, _p, _i, _i, _i64, _ip, _i, _i, _p, _i, _p, _i, _i, _i, _u64, _d,
                               _p, _p, _sz, _p]),
    "b200q_expval_csr": (_i, [_p, _i, _i, _i64, _ip, _i, _p, _p, _p, _p, _p, _sz, _p]),
    "b200q_apply_rtile_bcast": (_i, [_p, _p, _i, _i, _i64, _ip, _i, _i, _p, _i, _p, _i, _i, _i, _u64,
                                     _d, _p, _p, _sz, _p]),
    "b200q_jit_available": (_i, []),
    "b200q_jit_compile": (_i, [C.c_char_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), _i, _i, _i,
                               C.POINTER(_p), C.POINTER(_sz)]),
    "b200q_jit_free": (None, [_p]),
    "b200q_seg_load": (_i, [_p, _sz, C.POINTER(_p)]),
    "b200q_seg_unload": (_i, [_p]),
    "b200q_seg_launch": (_i, [_p, _p, _p, _i, _i, _i64, _ip, _i, _i, _i, _i, _ip, _i, _dp, _i, _i,
                              _i, _i, _u64, _u64, _u64, _d, _p, _p, _sz, _p]),
    "b200q_remap_copy": (_i, [_p, _sz, _p, _sz, _sz, _sz, _p]),
    "b200q_remap_unpack": (_i, [_p, _sz, _p, _sz, _sz, _sz, _i, _i, _p]),
    "b200q_stream_write32": (_i, [_p, C.c_uint32, _p]),
    "b200q_stream_wait_geq32": (_i, [_p, C.c_uint32, _p]),
    "b200q_adjoint_step": (_i, [_p, _i, _i, _i, _ip, _i, _ip, _ip, _i, _p, _p, _p, _p, _sz, _p]),
}


class B200QError(RuntimeError):
    """Raised when the native library is missing or a native call fails."""


@lru_cache(maxsize=1)
def load() -> C.CDLL:
    """Load ``libb200q.so`` (once).  Fails loudly when it has not been built."""
    if not os.path.exists(LIB_PATH):
        raise B200QError(
            f"native library not found at {LIB_PATH}: run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc). pennylane_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:  # pragma: no cover - build/ABI mismatch
            raise B200QError(f"{LIB_PATH} does not export {name}; rebuild it") from e
        fn.restype = res
        fn.argtypes = args
    return lib


def check(rc: int) -> None:
    """Turn a non-zero return code into an exception carrying ``b200q_last_error()``."""
    if rc != 0:
        msg = load().b200q_last_error()
        raise B200QError(f"b200q call failed (rc={rc}): {msg.decode() if msg else '?'}")


def int_array(values):
    """ctypes int array (or None for an empty list)."""
    values = list(values)
    if not values:
        return None
    return (C.c_int * len(values))(*[int(v) for v in values])


def u64_array(values):
    values = list(values)
    return (C.c_uint64 * max(1, len(values)))(*[int(v) for v in values])


def f64_array(values):
    values = list(values)
    return (C.c_double * max(1, len(values)))(*[float(v) for v in values])
