"""b200-qubit: a B200-native statevector simulator behind PennyLane's Device API.

``B200Qubit`` mirrors ``default.qubit`` (``execute`` / ``preprocess`` / ``compute_derivatives`` /
``compute_vjp`` ...) and runs every amplitude-sized operation as a hand-written sm_100a CUDA
kernel reached through the C ABI in ``include/b200q.h``.  With PennyLane installed the same
engine registers as ``qml.device("b200.qubit")`` (``pennylane_b200.pl_plugin``); without it the
operator / tape / measurement mirror classes in this package drive it directly.
"""
from . import mcm, measurements, one_shot, ops, pauli
from ._lib import B200QError, LIB_PATH
from .device import B200Qubit, DeviceError, ExecutionConfig, MCMConfig, QuantumFunctionError, device
from .mcm import cond, measure
from .measurements import (classical_shadow, counts, density_matrix, expval, mutual_info, probs, purity, sample,
                           shadow_expval, state, var, vn_entropy)
from .statevector import StateVector
from .tape import QuantumScript, QuantumTape, Shots

__version__ = "0.1.0"

__all__ = [
    "B200Qubit", "device", "ExecutionConfig", "MCMConfig", "DeviceError", "QuantumFunctionError", "StateVector",
    "QuantumScript", "QuantumTape", "Shots", "ops", "measurements", "pauli", "mcm", "one_shot",
    "measure", "cond",
    "expval", "var", "probs", "sample", "counts", "state", "density_matrix", "purity",
    "vn_entropy", "mutual_info", "classical_shadow", "shadow_expval", "B200QError", "LIB_PATH",
]
