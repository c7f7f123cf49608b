"""Operator classes for driving the engine without PennyLane installed.

This is a *mirror of the operator interface* the device consumes at its boundary
(``pennylane/core/operator/base.py``: ``name``, ``wires``, ``data``, ``parameters``,
``hyperparameters``, ``num_params``, ``batch_size``, ``matrix()``, ``has_matrix``,
``generator()``, ``has_generator``, ``diagonalizing_gates()``, ``eigvals()``, ``pauli_rep``,
``control_wires``, ``decomposition()``), with PennyLane's class names and argument meaning, so
that (a) the parity tests read like the reference's tests and (b) the device code is duck-typed
and accepts genuine ``pennylane`` operators unchanged when PennyLane is importable.

Matrix conventions follow the reference files cited per class.  Every ``compute_matrix``
accepts scalars or 1-D arrays (parameter broadcasting -> leading batch axis), like the
reference.
"""
from __future__ import annotations

import itertools
from typing import Sequence

import numpy as np

from .pauli import PauliSentence, PauliWord

INV_SQRT2 = 1 / np.sqrt(2)

_I2 = np.eye(2, dtype=complex)
_X = np.array([[0, 1], [1, 0]], dtype=complex)
_Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
_Z = np.array([[1, 0], [0, -1]], dtype=complex)
_H = np.array([[INV_SQRT2, INV_SQRT2], [INV_SQRT2, -INV_SQRT2]], dtype=complex)
_PAULI = {"I": _I2, "X": _X, "Y": _Y, "Z": _Z}


def _wires_tuple(wires) -> tuple:
    if wires is None:
        return ()
    if isinstance(wires, (str, bytes)) or not hasattr(wires, "__iter__"):
        return (wires,)
    return tuple(wires)


def _stack(rows):
    """rows: nested list of scalars / (B,) arrays -> (..., r, c) complex array."""
    rows = [[np.asarray(v, dtype=complex) for v in row] for row in rows]
    shape = np.broadcast(*[v for row in rows for v in row]).shape
    out = np.empty(shape + (len(rows), len(rows[0])), dtype=complex)
    for i, row in enumerate(rows):
        for j, v in enumerate(row):
            out[..., i, j] = v
    return out


def expand_matrix(mat: np.ndarray, wires: Sequence, wire_order: Sequence) -> np.ndarray:
    """Embed ``mat`` (acting on ``wires``, first wire = most significant) into ``wire_order``.
    Mirrors ``pennylane.math.expand_matrix`` (pennylane/math/matrix_manipulation.py)."""
    wires = list(wires)
    wire_order = list(wire_order)
    if wires == wire_order:
        return mat
    n, k = len(wire_order), len(wires)
    batch = mat.shape[:-2]
    extra = [w for w in wire_order if w not in wires]
    full = mat
    if extra:
        eye = np.eye(2 ** len(extra), dtype=mat.dtype)
        full = np.einsum("...ab,cd->...acbd", mat, eye).reshape(batch + (2**n, 2**n))
    cur = wires + extra
    perm = [cur.index(w) for w in wire_order]
    nb = len(batch)
    t = full.reshape(batch + (2,) * (2 * n))
    axes = list(range(nb)) + [nb + p for p in perm] + [nb + n + p for p in perm]
    return t.transpose(axes).reshape(batch + (2**n, 2**n))
    del k


class Operator:
    """Base operator.  ``ndim_params`` gives the un-broadcast rank of each parameter."""

    num_wires: int | None = None
    num_params: int = 0
    ndim_params: tuple = ()
    has_matrix = True
    has_generator = False
    has_diagonalizing_gates = False
    has_decomposition = False
    is_hermitian = False

    def __init__(self, *params, wires=None, id=None):
        self.wires = _wires_tuple(wires)
        if self.num_wires is not None and len(self.wires) != self.num_wires:
            raise ValueError(
                f"{self.name}: wrong number of wires. {len(self.wires)} wires given, "
                f"{self.num_wires} expected."
            )
        if len(set(self.wires)) != len(self.wires):
            raise ValueError(f"{self.name}: wires must be unique, got {self.wires}")
        if len(params) != self.num_params:
            raise ValueError(
                f"{self.name}: wrong number of parameters. {len(params)} parameters passed, "
                f"{self.num_params} expected."
            )
        self.data = tuple(params)
        self.id = id
        self.hyperparameters = {}

    # -- identification -----------------------------------------------------------------
    @property
    def name(self) -> str:
        return type(self).__name__

    @property
    def parameters(self) -> list:
        return list(self.data)

    @property
    def ndim_params_(self):
        return self.ndim_params or (0,) * self.num_params

    @property
    def batch_size(self):
        bs = None
        for p, nd in zip(self.data, self.ndim_params_):
            d = np.ndim(p)
            if d == nd + 1:
                b = np.shape(p)[0]
                if bs is not None and b != bs:
                    raise ValueError(f"{self.name}: inconsistent broadcasting dimensions")
                bs = b
            elif d != nd:
                raise ValueError(f"{self.name}: parameter with {d} dimensions, expected {nd}")
        return bs

    @property
    def control_wires(self):
        return ()

    def __repr__(self):
        ps = ", ".join(repr(p) if np.ndim(p) == 0 else f"<{np.shape(p)}>" for p in self.data)
        return f"{self.name}({ps}{', ' if ps else ''}wires={list(self.wires)})"

    # -- numerical representation --------------------------------------------------------
    @staticmethod
    def compute_matrix(*params, **hyper):  # pragma: no cover - abstract
        raise NotImplementedError

    def matrix(self, wire_order=None) -> np.ndarray:
        mat = self.compute_matrix(*self.data, **self.hyperparameters)
        if wire_order is None:
            return mat
        return expand_matrix(np.asarray(mat), self.wires, wire_order)

    def generator(self):  # pragma: no cover - abstract
        raise NotImplementedError(f"{self.name} has no generator")

    def diagonalizing_gates(self):
        raise NotImplementedError(f"{self.name} has no diagonalizing gates")

    def eigvals(self):
        return np.linalg.eigvals(self.matrix())

    def decomposition(self):
        raise NotImplementedError(f"{self.name} has no decomposition")

    @property
    def pauli_rep(self):
        return None

    def adjoint(self):
        return Adjoint(self)

    def map_wires(self, wire_map: dict):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new.wires = tuple(wire_map.get(w, w) for w in self.wires)
        new.hyperparameters = dict(self.hyperparameters)
        return new

    def _with_params(self, params):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        new.data = tuple(params)
        return new

    # -- observable arithmetic -------------------------------------------------------------
    def __matmul__(self, other):
        return Prod(self, other)

    def __mul__(self, s):
        return SProd(s, self)

    __rmul__ = __mul__

    def __add__(self, other):
        if isinstance(other, (int, float)) and other == 0:
            return self
        return Sum(self, other)

    __radd__ = __add__

    def __sub__(self, other):
        return Sum(self, SProd(-1.0, other))

    def __neg__(self):
        return SProd(-1.0, self)


Operation = Operator


# =============================================================================================
# Non-parametrised gates — pennylane/ops/qubit/non_parametric_ops.py, identity.py
# =============================================================================================
class Identity(Operator):
    """ops/identity.py:33."""
    num_params = 0
    is_hermitian = True
    has_diagonalizing_gates = True

    def __init__(self, wires=None, id=None):
        super().__init__(wires=wires, id=id)

    def matrix(self, wire_order=None):
        n = len(wire_order) if wire_order is not None else max(1, len(self.wires))
        return np.eye(2**n, dtype=complex)

    @staticmethod
    def compute_matrix(**_):
        return np.eye(2, dtype=complex)

    def diagonalizing_gates(self):
        return []

    def eigvals(self):
        return np.ones(2 ** max(1, len(self.wires)))

    @property
    def pauli_rep(self):
        return PauliSentence({PauliWord({}): 1.0})

    def adjoint(self):
        return Identity(wires=self.wires)


class _Fixed(Operator):
    _mat: np.ndarray = None

    def __init__(self, wires=None, id=None):
        super().__init__(wires=wires, id=id)

    @classmethod
    def compute_matrix(cls, **_):
        return cls._mat


class Hadamard(_Fixed):
    """non_parametric_ops.py:49 (matrix :98, eigvals :126, diagonalizing gates :152)."""
    num_wires = 1
    _mat = _H
    is_hermitian = True
    has_diagonalizing_gates = True

    def diagonalizing_gates(self):
        return [RY(-np.pi / 4, wires=self.wires)]

    def eigvals(self):
        return np.array([1.0, -1.0])

    def adjoint(self):
        return Hadamard(wires=self.wires)


class PauliX(_Fixed):
    """non_parametric_ops.py:296."""
    num_wires = 1
    _mat = _X
    is_hermitian = True
    has_diagonalizing_gates = True

    def diagonalizing_gates(self):
        return [Hadamard(wires=self.wires)]

    def eigvals(self):
        return np.array([1.0, -1.0])

    @property
    def pauli_rep(self):
        return PauliSentence({PauliWord({self.wires[0]: "X"}): 1.0})

    def adjoint(self):
        return PauliX(wires=self.wires)


class PauliY(_Fixed):
    """non_parametric_ops.py:512 (diagonalizing gates :628 = [Z, S, H])."""
    num_wires = 1
    _mat = _Y
    is_hermitian = True
    has_diagonalizing_gates = True

    def diagonalizing_gates(self):
        return [PauliZ(wires=self.wires), S(wires=self.wires), Hadamard(wires=self.wires)]

    def eigvals(self):
        return np.array([1.0, -1.0])

    @property
    def pauli_rep(self):
        return PauliSentence({PauliWord({self.wires[0]: "Y"}): 1.0})

    def adjoint(self):
        return PauliY(wires=self.wires)


class PauliZ(_Fixed):
    """non_parametric_ops.py:742."""
    num_wires = 1
    _mat = _Z
    is_hermitian = True
    has_diagonalizing_gates = True

    def diagonalizing_gates(self):
        return []

    def eigvals(self):
        return np.array([1.0, -1.0])

    @property
    def pauli_rep(self):
        return PauliSentence({PauliWord({self.wires[0]: "Z"}): 1.0})

    def adjoint(self):
        return PauliZ(wires=self.wires)


X, Y, Z, H = PauliX, PauliY, PauliZ, Hadamard


class S(_Fixed):
    """non_parametric_ops.py:1005."""
    num_wires = 1
    _mat = np.array([[1, 0], [0, 1j]], dtype=complex)


class T(_Fixed):
    """non_parametric_ops.py:1148."""
    num_wires = 1
    _mat = np.array([[1, 0], [0, np.exp(1j * np.pi / 4)]], dtype=complex)


class SX(_Fixed):
    """non_parametric_ops.py:1274."""
    num_wires = 1
    _mat = 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]], dtype=complex)


class SWAP(_Fixed):
    """non_parametric_ops.py:1406."""
    num_wires = 2
    _mat = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)

    def adjoint(self):
        return SWAP(wires=self.wires)


class ECR(_Fixed):
    """non_parametric_ops.py:1592."""
    num_wires = 2
    _mat = INV_SQRT2 * np.array(
        [[0, 0, 1, 1j], [0, 0, 1j, 1], [1, -1j, 0, 0], [-1j, 1, 0, 0]], dtype=complex)


class ISWAP(_Fixed):
    """non_parametric_ops.py:1730."""
    num_wires = 2
    _mat = np.array([[1, 0, 0, 0], [0, 0, 1j, 0], [0, 1j, 0, 0], [0, 0, 0, 1]], dtype=complex)


class SISWAP(_Fixed):
    """non_parametric_ops.py:1884."""
    num_wires = 2
    _mat = np.array(
        [[1, 0, 0, 0], [0, INV_SQRT2, INV_SQRT2 * 1j, 0], [0, INV_SQRT2 * 1j, INV_SQRT2, 0],
         [0, 0, 0, 1]], dtype=complex)


SQISW = SISWAP


def _controlled_mat(base: np.ndarray, n_ctrl: int, control_values=None) -> np.ndarray:
    """Block matrix: ``base`` acts when the controls match ``control_values`` (default all 1).
    op_math/controlled.py (Controlled.matrix)."""
    base = np.asarray(base)
    batch = base.shape[:-2]
    d = base.shape[-1]
    D = d * 2**n_ctrl
    vals = [1] * n_ctrl if control_values is None else [int(bool(v)) for v in control_values]
    sel = int("".join(str(v) for v in vals), 2) if n_ctrl else 0
    out = np.zeros(batch + (D, D), dtype=complex)
    out[..., np.arange(D), np.arange(D)] = 1.0
    lo = sel * d
    out[..., lo:lo + d, lo:lo + d] = base
    return out


class CNOT(_Fixed):
    """op_math/controlled_ops.py:945."""
    num_wires = 2
    _mat = _controlled_mat(_X, 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def adjoint(self):
        return CNOT(wires=self.wires)


class CZ(_Fixed):
    """op_math/controlled_ops.py:522."""
    num_wires = 2
    _mat = _controlled_mat(_Z, 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def adjoint(self):
        return CZ(wires=self.wires)


class CY(_Fixed):
    """op_math/controlled_ops.py:378."""
    num_wires = 2
    _mat = _controlled_mat(_Y, 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def adjoint(self):
        return CY(wires=self.wires)


class CH(_Fixed):
    """op_math/controlled_ops.py:279."""
    num_wires = 2
    _mat = _controlled_mat(_H, 1)

    @property
    def control_wires(self):
        return self.wires[:1]


class CSWAP(_Fixed):
    """op_math/controlled_ops.py:648."""
    num_wires = 3
    _mat = _controlled_mat(SWAP._mat, 1)

    @property
    def control_wires(self):
        return self.wires[:1]


class Toffoli(_Fixed):
    """op_math/controlled_ops.py:1087."""
    num_wires = 3
    _mat = _controlled_mat(_X, 2)

    @property
    def control_wires(self):
        return self.wires[:2]

    def adjoint(self):
        return Toffoli(wires=self.wires)


class CCZ(_Fixed):
    """op_math/controlled_ops.py:782."""
    num_wires = 3
    _mat = _controlled_mat(_Z, 2)

    @property
    def control_wires(self):
        return self.wires[:2]


class MultiControlledX(Operator):
    """op_math/controlled_ops.py:1284.  ``wires`` = controls + [target]."""

    def __init__(self, wires=None, control_values=None, id=None):
        super().__init__(wires=wires, id=id)
        if len(self.wires) < 1:
            raise ValueError("MultiControlledX needs at least one wire")
        nc = len(self.wires) - 1
        if control_values is None:
            control_values = [1] * nc
        if isinstance(control_values, str):
            control_values = [int(c) for c in control_values]
        if len(control_values) != nc:
            raise ValueError("control_values must match the number of control wires")
        self.hyperparameters["control_values"] = [bool(v) for v in control_values]

    @property
    def control_wires(self):
        return self.wires[:-1]

    @property
    def control_values(self):
        return self.hyperparameters["control_values"]

    def matrix(self, wire_order=None):
        mat = _controlled_mat(_X, len(self.wires) - 1, self.control_values)
        return mat if wire_order is None else expand_matrix(mat, self.wires, wire_order)

    def adjoint(self):
        return MultiControlledX(wires=self.wires, control_values=self.control_values)


class GroverOperator(Operator):
    """templates/subroutines/grover.py:29 — the diffusion operator ``G = 2|s><s| - I`` on
    ``wires`` (|s> = uniform superposition).  The reference applies it matrix-free on nine or
    more wires (apply_operation.py:836-880: sum over the operator's axes, refill with the
    all-plus state); here those widths go through the decomposition
    ``H^k (2|0><0| - I) H^k`` — Hadamards that fuse into neighbouring blocks, one multiply
    controlled phase that touches only 2^-k of the state, and a global phase."""

    has_decomposition = True

    def __init__(self, wires=None, work_wires=None, id=None):
        super().__init__(wires=wires, id=id)
        if len(self.wires) < 2:
            raise ValueError("GroverOperator must have at least two wires. "
                             f"Got {len(self.wires)} wires.")
        self.hyperparameters["n_wires"] = len(self.wires)
        self.hyperparameters["work_wires"] = _wires_tuple(work_wires) if work_wires is not None else ()

    def matrix(self, wire_order=None):
        dim = 1 << len(self.wires)
        mat = np.full((dim, dim), 2.0 / dim, dtype=complex) - np.eye(dim)
        return mat if wire_order is None else expand_matrix(mat, self.wires, wire_order)

    def decomposition(self):
        w = self.wires
        t = w[-1]
        flip0 = [PauliX(wires=t), Controlled(PauliZ(wires=t), control_wires=w[:-1],
                                             control_values=[False] * (len(w) - 1)), PauliX(wires=t)]
        had = [Hadamard(wires=x) for x in w]
        # 2|0><0| - I = -(I - 2|0><0|): the flip of |0..0> followed by a global -1 = exp(-i pi)
        return had + flip0 + had + [GlobalPhase(np.pi, wires=w)]

    def adjoint(self):
        return GroverOperator(wires=self.wires)


# =============================================================================================
# Parametrised single-qubit gates — pennylane/ops/qubit/parametric_ops_single_qubit.py
# =============================================================================================
class _OneParam(Operator):
    num_params = 1

    def __init__(self, phi, wires=None, id=None):
        super().__init__(phi, wires=wires, id=id)

    def adjoint(self):
        return self._with_params([-np.asarray(self.data[0]) if np.ndim(self.data[0]) else -self.data[0]])


class RX(_OneParam):
    """:62 — matrix :117-148 ``[[c, -is], [-is, c]]``, generator :111 ``-0.5 X``."""
    num_wires = 1
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        c, s = np.cos(np.asarray(phi) / 2), np.sin(np.asarray(phi) / 2)
        return _stack([[c, -1j * s], [-1j * s, c]])

    def generator(self):
        return SProd(-0.5, PauliX(wires=self.wires))


class RY(_OneParam):
    """:273 — matrix :328 ``[[c, -s], [s, c]]``, generator :322 ``-0.5 Y``."""
    num_wires = 1
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        c, s = np.cos(np.asarray(phi) / 2), np.sin(np.asarray(phi) / 2)
        return _stack([[c, -s], [s, c]])

    def generator(self):
        return SProd(-0.5, PauliY(wires=self.wires))


class RZ(_OneParam):
    """:497 — matrix :548-578 ``diag(e^{-i phi/2}, e^{i phi/2})``, generator :540 ``-0.5 Z``."""
    num_wires = 1
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        z = np.zeros_like(phi, dtype=complex)
        return _stack([[np.exp(-0.5j * phi), z], [z, np.exp(0.5j * phi)]])

    def generator(self):
        return SProd(-0.5, PauliZ(wires=self.wires))


class PhaseShift(_OneParam):
    """:780 — matrix :839 ``diag(1, e^{i phi})``, generator :826 ``Projector([1])``."""
    num_wires = 1
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        z = np.zeros_like(phi, dtype=complex)
        return _stack([[z + 1, z], [z, np.exp(1j * phi)]])

    def generator(self):
        return Projector(np.array([1]), wires=self.wires)


class U1(PhaseShift):
    """:1193 — same matrix as PhaseShift (:1238)."""


class Rot(Operator):
    """:989 — ``RZ(omega) RY(theta) RZ(phi)``, matrix :1036-1095, decomposition RZ,RY,RZ."""
    num_wires = 1
    num_params = 3
    has_decomposition = True

    def __init__(self, phi, theta, omega, wires=None, id=None):
        super().__init__(phi, theta, omega, wires=wires, id=id)

    @staticmethod
    def compute_matrix(phi, theta, omega, **_):
        phi, theta, omega = np.asarray(phi), np.asarray(theta), np.asarray(omega)
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        return _stack([
            [np.exp(-0.5j * (phi + omega)) * c, -np.exp(0.5j * (phi - omega)) * s],
            [np.exp(-0.5j * (phi - omega)) * s, np.exp(0.5j * (phi + omega)) * c],
        ])

    def decomposition(self):
        phi, theta, omega = self.data
        return [RZ(phi, wires=self.wires), RY(theta, wires=self.wires), RZ(omega, wires=self.wires)]

    def adjoint(self):
        phi, theta, omega = self.data
        return Rot(-omega, -theta, -phi, wires=self.wires)


class U2(Operator):
    """:1312 — matrix :1364-1400; decomposition Rot(delta,pi/2,-delta) PhaseShift(delta)
    PhaseShift(phi) (:1402-1440)."""
    num_wires = 1
    num_params = 2
    has_decomposition = True

    def __init__(self, phi, delta, wires=None, id=None):
        super().__init__(phi, delta, wires=wires, id=id)

    @staticmethod
    def compute_matrix(phi, delta, **_):
        phi, delta = np.asarray(phi), np.asarray(delta)
        one = np.ones(np.broadcast(phi, delta).shape, dtype=complex)
        return INV_SQRT2 * _stack([
            [one, -np.exp(1j * delta) * one],
            [np.exp(1j * phi) * one, np.exp(1j * (phi + delta))],
        ])

    def decomposition(self):
        phi, delta = self.data
        w = self.wires
        return [Rot(delta, np.pi / 2, -np.asarray(delta) if np.ndim(delta) else -delta, wires=w),
                PhaseShift(delta, wires=w), PhaseShift(phi, wires=w)]


class U3(Operator):
    """:1450 — matrix :1509-1560; decomposition Rot(delta,theta,-delta) PhaseShift(delta)
    PhaseShift(phi)."""
    num_wires = 1
    num_params = 3
    has_decomposition = True

    def __init__(self, theta, phi, delta, wires=None, id=None):
        super().__init__(theta, phi, delta, wires=wires, id=id)

    @staticmethod
    def compute_matrix(theta, phi, delta, **_):
        theta, phi, delta = np.asarray(theta), np.asarray(phi), np.asarray(delta)
        c, s = np.cos(theta / 2), np.sin(theta / 2)
        one = np.ones(np.broadcast(theta, phi, delta).shape, dtype=complex)
        return _stack([
            [c * one, -s * np.exp(1j * delta) * one],
            [s * np.exp(1j * phi) * one, c * np.exp(1j * (phi + delta))],
        ])

    def decomposition(self):
        theta, phi, delta = self.data
        w = self.wires
        return [Rot(delta, theta, -np.asarray(delta) if np.ndim(delta) else -delta, wires=w),
                PhaseShift(delta, wires=w), PhaseShift(phi, wires=w)]


class GlobalPhase(_OneParam):
    """ops/identity.py:235 (GlobalPhase): ``exp(-i phi)`` on the whole state
    (devices/qubit/apply_operation.py:507-517); generator ``-Identity``."""
    has_generator = True

    def __init__(self, phi, wires=None, id=None):
        super().__init__(phi, wires=wires, id=id)

    def matrix(self, wire_order=None):
        n = len(wire_order) if wire_order is not None else max(1, len(self.wires))
        ph = np.exp(-1j * np.asarray(self.data[0]))
        eye = np.eye(2**n, dtype=complex)
        return ph[..., None, None] * eye if np.ndim(ph) else ph * eye

    def generator(self):
        return SProd(-1.0, Identity(wires=self.wires))


# =============================================================================================
# Parametrised multi-qubit gates — pennylane/ops/qubit/parametric_ops_multi_qubit.py
# =============================================================================================
def _pauli_string_matrix(word: str) -> np.ndarray:
    m = np.array([[1.0 + 0j]])
    for ch in word:
        m = np.kron(m, _PAULI[ch])
    return m


def _exp_pauli(theta, P: np.ndarray) -> np.ndarray:
    """exp(-i theta/2 P) for an involutory P."""
    theta = np.asarray(theta)
    c, s = np.cos(theta / 2), np.sin(theta / 2)
    eye = np.eye(P.shape[0], dtype=complex)
    if theta.ndim:
        return c[:, None, None] * eye - 1j * s[:, None, None] * P
    return c * eye - 1j * s * P


class MultiRZ(_OneParam):
    """:47 — ``exp(-i theta/2 Z^{(x)n})``, matrix :93, generator :133 ``-0.5 Z..Z``."""
    has_generator = True

    def __init__(self, theta, wires=None, id=None):
        super().__init__(theta, wires=wires, id=id)
        self.hyperparameters["num_wires"] = len(self.wires)

    @staticmethod
    def compute_matrix(theta, num_wires=1, **_):
        return _exp_pauli(theta, _pauli_string_matrix("Z" * num_wires))

    def generator(self):
        return SProd(-0.5, Prod(*[PauliZ(wires=w) for w in self.wires]))


class PauliRot(_OneParam):
    """:229 — ``exp(-i theta/2 P)``, matrix :380-436, generator :451 ``-0.5 P``."""
    has_generator = True

    def __init__(self, theta, pauli_word, wires=None, id=None):
        super().__init__(theta, wires=wires, id=id)
        if len(pauli_word) != len(self.wires) or any(c not in "IXYZ" for c in pauli_word):
            raise ValueError(
                f'The given Pauli word "{pauli_word}" contains characters that are not allowed '
                "or has the wrong length. Allowed characters are I, X, Y and Z"
            )
        self.hyperparameters["pauli_word"] = pauli_word

    @staticmethod
    def compute_matrix(theta, pauli_word="", **_):
        return _exp_pauli(theta, _pauli_string_matrix(pauli_word))

    def generator(self):
        word = self.hyperparameters["pauli_word"]
        return SProd(-0.5, pauli_word_op(word, self.wires))


def pauli_word_op(word: str, wires):
    ops_ = [{"X": PauliX, "Y": PauliY, "Z": PauliZ, "I": Identity}[c](wires=w)
            for c, w in zip(word, wires)]
    return ops_[0] if len(ops_) == 1 else Prod(*ops_)


class IsingXX(_OneParam):
    """:1071 — matrix :1121, generator :1113 ``-0.5 XX``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _exp_pauli(phi, _pauli_string_matrix("XX"))

    def generator(self):
        return SProd(-0.5, PauliX(wires=self.wires[0]) @ PauliX(wires=self.wires[1]))


class IsingYY(_OneParam):
    """:1206 — matrix :1256, generator :1248 ``-0.5 YY``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _exp_pauli(phi, _pauli_string_matrix("YY"))

    def generator(self):
        return SProd(-0.5, PauliY(wires=self.wires[0]) @ PauliY(wires=self.wires[1]))


class IsingZZ(_OneParam):
    """:1349 — matrix :1400, generator :1392 ``-0.5 ZZ``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _exp_pauli(phi, _pauli_string_matrix("ZZ"))

    def generator(self):
        return SProd(-0.5, PauliZ(wires=self.wires[0]) @ PauliZ(wires=self.wires[1]))


class IsingXY(_OneParam):
    """:1525 — matrix :1592-1645 ``diag(1,c,c,1) + i s (|01><10| + |10><01|)``,
    generator :1578 ``0.25 (XX + YY)``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        z = np.zeros_like(phi, dtype=complex)
        o = z + 1
        return _stack([[o, z, z, z], [z, c, 1j * s, z], [z, 1j * s, c, z], [z, z, z, o]])

    def generator(self):
        w0, w1 = self.wires
        return LinearCombination(
            [0.25, 0.25], [PauliX(wires=w0) @ PauliX(wires=w1), PauliY(wires=w0) @ PauliY(wires=w1)])


class PSWAP(_OneParam):
    """:1718 — matrix :1791-1836."""
    num_wires = 2

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        e = np.exp(1j * phi)
        z = np.zeros_like(phi, dtype=complex)
        o = z + 1
        return _stack([[o, z, z, z], [z, z, e, z], [z, e, z, z], [z, z, z, o]])


class CRX(_OneParam):
    """op_math/controlled_ops.py:1546; generator ``-0.5 |1><1| (x) X``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _controlled_mat(RX.compute_matrix(phi), 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def generator(self):
        return SProd(-0.5, Projector(np.array([1]), wires=self.wires[0]) @ PauliX(wires=self.wires[1]))


class CRY(_OneParam):
    """op_math/controlled_ops.py:1715."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _controlled_mat(RY.compute_matrix(phi), 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def generator(self):
        return SProd(-0.5, Projector(np.array([1]), wires=self.wires[0]) @ PauliY(wires=self.wires[1]))


class CRZ(_OneParam):
    """op_math/controlled_ops.py:1859."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _controlled_mat(RZ.compute_matrix(phi), 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def generator(self):
        return SProd(-0.5, Projector(np.array([1]), wires=self.wires[0]) @ PauliZ(wires=self.wires[1]))


class CRot(Operator):
    """op_math/controlled_ops.py:2052; decomposition :2170-2208
    (RZ, RY, CNOT, RY, RZ, CNOT, RZ)."""
    num_wires = 2
    num_params = 3
    has_decomposition = True

    def __init__(self, phi, theta, omega, wires=None, id=None):
        super().__init__(phi, theta, omega, wires=wires, id=id)

    @staticmethod
    def compute_matrix(phi, theta, omega, **_):
        return _controlled_mat(Rot.compute_matrix(phi, theta, omega), 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def decomposition(self):
        phi, theta, omega = (np.asarray(p) if np.ndim(p) else p for p in self.data)
        c, t = self.wires
        return [
            RZ((phi - omega) / 2, wires=t), CNOT(wires=[c, t]),
            RZ(-(phi + omega) / 2, wires=t), RY(-theta / 2, wires=t), CNOT(wires=[c, t]),
            RY(theta / 2, wires=t), RZ(omega, wires=t),
        ]

    def adjoint(self):
        phi, theta, omega = self.data
        return CRot(-omega, -theta, -phi, wires=self.wires)


class ControlledPhaseShift(_OneParam):
    """op_math/controlled_ops.py:2210; generator ``|11><11|``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        return _controlled_mat(PhaseShift.compute_matrix(phi), 1)

    @property
    def control_wires(self):
        return self.wires[:1]

    def generator(self):
        return Projector(np.array([1, 1]), wires=self.wires)


CPhase = ControlledPhaseShift


# ---- qchem gates: pennylane/ops/qubit/qchem_ops.py ---------------------------------------------
class SingleExcitation(_OneParam):
    """qchem_ops.py:124 — matrix via ``_single_excitations_matrix`` (:40-88), generator :180
    ``0.25 (X0 Y1 - Y0 X1)``."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        z = np.zeros_like(phi, dtype=complex)
        o = z + 1
        return _stack([[o, z, z, z], [z, c, -s, z], [z, s, c, z], [z, z, z, o]])

    def generator(self):
        w0, w1 = self.wires
        return LinearCombination(
            [0.25, -0.25], [PauliX(wires=w0) @ PauliY(wires=w1), PauliY(wires=w0) @ PauliX(wires=w1)])


class SingleExcitationMinus(_OneParam):
    """qchem_ops.py:270 — phase ``e^{-i phi/2}`` outside the excitation subspace, generator :316."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        e = np.exp(-0.5j * phi)
        z = np.zeros_like(phi, dtype=complex)
        return _stack([[e, z, z, z], [z, c, -s, z], [z, s, c, z], [z, z, z, e]])

    def generator(self):
        w0, w1 = self.wires
        return LinearCombination(
            [-0.25, 0.25, -0.25, -0.25],
            [Identity(wires=w0), PauliX(wires=w0) @ PauliY(wires=w1),
             PauliY(wires=w0) @ PauliX(wires=w1), PauliZ(wires=w0) @ PauliZ(wires=w1)])


class SingleExcitationPlus(_OneParam):
    """qchem_ops.py:441 — phase ``e^{+i phi/2}`` outside the excitation subspace, generator :487."""
    num_wires = 2
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        e = np.exp(0.5j * phi)
        z = np.zeros_like(phi, dtype=complex)
        return _stack([[e, z, z, z], [z, c, -s, z], [z, s, c, z], [z, z, z, e]])

    def generator(self):
        w0, w1 = self.wires
        return LinearCombination(
            [0.25, 0.25, -0.25, 0.25],
            [Identity(wires=w0), PauliX(wires=w0) @ PauliY(wires=w1),
             PauliY(wires=w0) @ PauliX(wires=w1), PauliZ(wires=w0) @ PauliZ(wires=w1)])


class DoubleExcitation(_OneParam):
    """qchem_ops.py:605 — rotation in span{|0011>, |1100>} (:90-120, mask :697-699),
    generator :675."""
    num_wires = 4
    has_generator = True

    @staticmethod
    def compute_matrix(phi, **_):
        phi = np.asarray(phi)
        c, s = np.cos(phi / 2), np.sin(phi / 2)
        mat = np.zeros(phi.shape + (16, 16), dtype=complex)
        mat[..., np.arange(16), np.arange(16)] = 1.0
        mat[..., 3, 3] = c
        mat[..., 12, 12] = c
        mat[..., 3, 12] = -s
        mat[..., 12, 3] = s
        return mat

    def generator(self):
        w = self.wires
        cs = [0.0625, 0.0625, -0.0625, 0.0625, -0.0625, 0.0625, -0.0625, -0.0625]
        words = ["XXXY", "XXYX", "XYXX", "XYYY", "YXXX", "YXYY", "YYXY", "YYYX"]
        return LinearCombination(cs, [pauli_word_op(wd, w) for wd in words])


# =============================================================================================
# Matrix-defined operators — pennylane/ops/qubit/matrix_ops.py, observables.py
# =============================================================================================
class QubitUnitary(Operator):
    """matrix_ops.py:91."""
    num_params = 1
    ndim_params = (2,)

    def __init__(self, U, wires=None, id=None, unitary_check=False):
        U = np.asarray(U)
        super().__init__(U, wires=wires, id=id)
        dim = 2 ** len(self.wires)
        if U.shape[-2:] != (dim, dim) or U.ndim not in (2, 3):
            raise ValueError(
                f"Input unitary must be of shape {(dim, dim)} or (batch_size, {dim}, {dim}) "
                f"to act on {len(self.wires)} wires. Got shape {U.shape} instead."
            )
        if unitary_check and not np.allclose(U @ np.conj(np.swapaxes(U, -1, -2)), np.eye(dim)):
            raise ValueError("Operator must be unitary.")

    @staticmethod
    def compute_matrix(U, **_):
        return np.asarray(U, dtype=complex)

    def adjoint(self):
        U = np.asarray(self.data[0])
        return QubitUnitary(np.conj(np.swapaxes(U, -1, -2)), wires=self.wires)


class DiagonalQubitUnitary(Operator):
    """matrix_ops.py:452."""
    num_params = 1
    ndim_params = (1,)

    def __init__(self, D, wires=None, id=None):
        D = np.asarray(D)
        super().__init__(D, wires=wires, id=id)
        if D.shape[-1] != 2 ** len(self.wires):
            raise ValueError("DiagonalQubitUnitary: wrong diagonal length")

    @staticmethod
    def compute_matrix(D, **_):
        D = np.asarray(D, dtype=complex)
        out = np.zeros(D.shape + (D.shape[-1],), dtype=complex)
        idx = np.arange(D.shape[-1])
        out[..., idx, idx] = D
        return out

    def eigvals(self):
        return np.asarray(self.data[0], dtype=complex)

    def adjoint(self):
        return DiagonalQubitUnitary(np.conj(self.data[0]), wires=self.wires)


class Hermitian(Operator):
    """observables.py:34 — observable given by a Hermitian matrix; diagonalizing gates
    ``QubitUnitary(eigvecs^dagger)`` (:170-190)."""
    num_params = 1
    ndim_params = (2,)
    is_hermitian = True
    has_diagonalizing_gates = True

    def __init__(self, A, wires=None, id=None):
        A = np.asarray(A)
        super().__init__(A, wires=wires, id=id)
        dim = 2 ** len(self.wires)
        if A.shape != (dim, dim):
            raise ValueError(f"Observable must be of shape {(dim, dim)}, got {A.shape}")
        if not np.allclose(A, A.conj().T):
            raise ValueError("Observable must be Hermitian.")
        self._eig = None

    @staticmethod
    def compute_matrix(A, **_):
        return np.asarray(A, dtype=complex)

    def _eigendecomposition(self):
        if self._eig is None:
            w, v = np.linalg.eigh(np.asarray(self.data[0], dtype=complex))
            self._eig = (w, v)
        return self._eig

    def eigvals(self):
        return self._eigendecomposition()[0]

    def diagonalizing_gates(self):
        v = self._eigendecomposition()[1]
        return [QubitUnitary(v.conj().T, wires=self.wires)]


class SparseHamiltonian(Operator):
    """ops/qubit/observables.py:279 — an observable given as a scipy CSR matrix on ``wires``
    (measured through the CSR reduction, measure.py:74-118; no dense matrix, no Pauli form)."""

    has_matrix = False
    is_hermitian = True
    num_params = 1

    def __init__(self, H, wires=None, id=None):
        import scipy.sparse as sp

        if not sp.issparse(H):
            raise TypeError("Observable must be a scipy sparse csr_matrix.")
        self.wires = _wires_tuple(wires)
        H = sp.csr_matrix(H)
        if H.shape != (1 << len(self.wires),) * 2:
            raise ValueError(f"Sparse Matrix must be of shape {(1 << len(self.wires),) * 2}.")
        self.data = (H,)
        self.id = id
        self.hyperparameters = {}

    name = "SparseHamiltonian"
    batch_size = None
    pauli_rep = None

    def sparse_matrix(self, wire_order=None):
        import scipy.sparse as sp

        H = self.data[0]
        if wire_order is None or list(wire_order) == list(self.wires):
            return H
        # observables.py:330-336: kron with identities, then permute to wire_order
        wire_order = list(wire_order)
        extra = [w for w in wire_order if w not in self.wires]
        full = sp.kron(H, sp.identity(1 << len(extra), format="csr"), format="csr") if extra else H
        cur = list(self.wires) + extra
        n = len(cur)
        idx = np.arange(1 << n)
        perm = np.zeros_like(idx)
        for pos, w in enumerate(wire_order):              # bit of w in the new order <- old order
            old = cur.index(w)
            perm |= ((idx >> (n - 1 - pos)) & 1) << (n - 1 - old)
        return full[perm][:, perm].tocsr()

    def matrix(self, wire_order=None):
        return np.asarray(self.sparse_matrix(wire_order).toarray())


class Projector(Operator):
    """observables.py:412 — ``|b><b|`` for a basis-state bit string (or a state vector)."""
    num_params = 1
    ndim_params = (1,)
    is_hermitian = True
    has_diagonalizing_gates = True

    def __init__(self, state, wires=None, id=None):
        state = np.asarray(state)
        super().__init__(state, wires=wires, id=id)
        n = len(self.wires)
        if state.shape == (n,):
            self._basis = True
        elif state.shape == (2**n,):
            self._basis = False
        else:
            raise ValueError(f"Input state must be of length {n} or {2**n}; got {state.shape}")

    def matrix(self, wire_order=None):
        n = len(self.wires)
        if self._basis:
            idx = int("".join(str(int(b)) for b in self.data[0]), 2)
            m = np.zeros((2**n, 2**n), dtype=complex)
            m[idx, idx] = 1.0
        else:
            v = np.asarray(self.data[0], dtype=complex)
            m = np.outer(v, v.conj())
        return m if wire_order is None else expand_matrix(m, self.wires, wire_order)

    def eigvals(self):
        n = len(self.wires)
        if self._basis:
            idx = int("".join(str(int(b)) for b in self.data[0]), 2)
            w = np.zeros(2**n)
            w[idx] = 1.0
            return w
        w = np.zeros(2**n)
        w[-1] = 1.0
        return w

    def diagonalizing_gates(self):
        if self._basis:
            return []
        v = np.asarray(self.data[0], dtype=complex)
        # unitary whose last row is <v| (eigenvalue 1 last), rest an orthonormal completion
        q, _ = np.linalg.qr(np.column_stack([v, np.eye(len(v), dtype=complex)[:, : len(v) - 1]]))
        q[:, 0] = v / np.linalg.norm(v)
        # Gram-Schmidt the rest against v
        basis = [q[:, 0]]
        for k in range(len(v)):
            e = np.eye(len(v), dtype=complex)[:, k]
            for b in basis:
                e = e - np.vdot(b, e) * b
            nrm = np.linalg.norm(e)
            if nrm > 1e-10:
                basis.append(e / nrm)
            if len(basis) == len(v):
                break
        U = np.array(basis[1:] + basis[:1]).conj()
        return [QubitUnitary(U, wires=self.wires)]


# =============================================================================================
# State preparation — pennylane/ops/qubit/state_preparation.py
# =============================================================================================
class StatePrepBase(Operator):
    has_matrix = False

    def state_vector(self, wire_order=None):  # pragma: no cover - abstract
        raise NotImplementedError


class BasisState(StatePrepBase):
    """state_preparation.py:43."""
    num_params = 1
    ndim_params = (1,)

    def __init__(self, state, wires=None, id=None):
        state = np.asarray(state)
        super().__init__(state, wires=wires, id=id)
        if state.shape != (len(self.wires),):
            raise ValueError(
                f"State must be of length {len(self.wires)}; got length {state.shape[-1]}")
        if not set(state.tolist()).issubset({0, 1}):
            raise ValueError(f"Basis state must only consist of 0s and 1s; got {state.tolist()}")

    def state_vector(self, wire_order=None):
        wire_order = self.wires if wire_order is None else tuple(wire_order)
        if not set(self.wires).issubset(wire_order):
            raise ValueError("wire_order must contain all BasisState wires")
        n = len(wire_order)
        idx = 0
        for w, b in zip(self.wires, self.data[0]):
            idx |= int(b) << (n - 1 - wire_order.index(w))
        out = np.zeros(2**n, dtype=complex)
        out[idx] = 1.0
        return out.reshape((2,) * n)


class StatePrep(StatePrepBase):
    """state_preparation.py:192 (``state_vector`` :420-470)."""
    num_params = 1
    ndim_params = (1,)

    def __init__(self, state, wires=None, pad_with=None, normalize=False, id=None, validate_norm=True):
        state = np.asarray(state)
        wires_t = _wires_tuple(wires)
        dim = 2 ** len(wires_t)
        if pad_with is not None and state.shape[-1] < dim:
            pad = np.full(state.shape[:-1] + (dim - state.shape[-1],), pad_with, dtype=state.dtype)
            state = np.concatenate([state, pad], axis=-1)
        if normalize:
            state = state / np.linalg.norm(state, axis=-1, keepdims=True)
        super().__init__(state, wires=wires, id=id)
        if state.ndim not in (1, 2) or state.shape[-1] != dim:
            raise ValueError(
                f"State must be of length {dim}; got length {state.shape[-1]}. Use the 'pad_with' "
                "argument for automated padding.")
        if validate_norm:
            nrm = np.linalg.norm(state, axis=-1)
            if not np.allclose(nrm, 1.0, atol=1e-10):
                raise ValueError("The state must be a vector of norm 1.0; got norm "
                                 f"{nrm}. Use 'normalize=True' to automatically normalize.")

    def state_vector(self, wire_order=None):
        st = np.asarray(self.data[0])
        bs = st.shape[0] if st.ndim == 2 else None
        k = len(self.wires)
        shape = ((bs,) if bs else ()) + (2,) * k
        st = st.reshape(shape)
        if wire_order is None or tuple(wire_order) == self.wires:
            return st
        wire_order = tuple(wire_order)
        if not set(self.wires).issubset(wire_order):
            raise ValueError("Custom wire_order must contain all StatePrep wires")
        extra = [w for w in wire_order if w not in self.wires]
        nb = 1 if bs else 0
        for _ in extra:
            st = np.stack([st, np.zeros_like(st)], axis=-1)
        cur = list(self.wires) + extra
        perm = [cur.index(w) for w in wire_order]
        return st.transpose(list(range(nb)) + [nb + p for p in perm])


# =============================================================================================
# Operator arithmetic — pennylane/ops/op_math/
# =============================================================================================
class Adjoint(Operator):
    """op_math/adjoint.py:290 (matrix = conjugate transpose, :396-398)."""

    def __init__(self, base: Operator, id=None):
        self.base = base
        self.wires = base.wires
        self.data = base.data
        self.id = id
        self.hyperparameters = {"base": base}

    num_params = property(lambda self: self.base.num_params)
    ndim_params = property(lambda self: self.base.ndim_params)
    has_generator = property(lambda self: self.base.has_generator)

    @property
    def name(self):
        return f"Adjoint({self.base.name})"

    @property
    def batch_size(self):
        return self.base.batch_size

    @property
    def control_wires(self):
        return self.base.control_wires

    def matrix(self, wire_order=None):
        m = np.asarray(self.base.matrix(wire_order=wire_order))
        return np.conj(np.swapaxes(m, -1, -2))

    def generator(self):
        return SProd(-1.0, self.base.generator())

    def adjoint(self):
        return self.base

    def map_wires(self, wire_map):
        return Adjoint(self.base.map_wires(wire_map))

    def _with_params(self, params):
        return Adjoint(self.base._with_params(params))


def adjoint(op: Operator) -> Operator:
    """``qml.adjoint(op)`` for an instantiated operator: uses the op's own adjoint rule when it
    has one (op_math/adjoint.py:120-150, lazy=False semantics), else wraps it."""
    return op.adjoint()


class Controlled(Operator):
    """op_math/controlled.py:486 — ``base`` applied when ``control_wires`` match ``control_values``."""

    def __init__(self, base: Operator, control_wires, control_values=None, id=None):
        self.base = base
        cw = _wires_tuple(control_wires)
        if set(cw) & set(base.wires):
            raise ValueError("The control wires must be different from the base operation wires.")
        if control_values is None:
            control_values = [True] * len(cw)
        if isinstance(control_values, (int, bool)):
            control_values = [control_values]
        if len(control_values) != len(cw):
            raise ValueError("control_values should be the same length as control_wires")
        self._control_wires = cw
        self.wires = cw + tuple(base.wires)
        self.data = base.data
        self.id = id
        self.hyperparameters = {"control_wires": cw,
                                "control_values": [bool(v) for v in control_values],
                                "base": base}

    num_params = property(lambda self: self.base.num_params)
    ndim_params = property(lambda self: self.base.ndim_params)

    @property
    def name(self):
        return f"C({self.base.name})"

    @property
    def batch_size(self):
        return self.base.batch_size

    @property
    def control_wires(self):
        return self._control_wires

    @property
    def control_values(self):
        return self.hyperparameters["control_values"]

    @property
    def has_generator(self):
        return self.base.has_generator

    def matrix(self, wire_order=None):
        m = _controlled_mat(np.asarray(self.base.matrix()), len(self._control_wires),
                            self.control_values)
        return m if wire_order is None else expand_matrix(m, self.wires, wire_order)

    def generator(self):
        proj = Projector(np.array([int(v) for v in self.control_values]), wires=self._control_wires)
        return Prod(proj, self.base.generator())

    def adjoint(self):
        return Controlled(self.base.adjoint(), self._control_wires, self.control_values)

    def map_wires(self, wire_map):
        return Controlled(self.base.map_wires(wire_map),
                          [wire_map.get(w, w) for w in self._control_wires], self.control_values)

    def _with_params(self, params):
        return Controlled(self.base._with_params(params), self._control_wires, self.control_values)


def ctrl(op: Operator, control, control_values=None) -> Operator:
    """``qml.ctrl(op, control, control_values)`` for an instantiated operator."""
    return Controlled(op, control, control_values)


class ControlledQubitUnitary(Controlled):
    """op_math/controlled_ops.py:80."""

    def __init__(self, U, wires=None, control_values=None, id=None):
        wires = _wires_tuple(wires)
        U = np.asarray(U)
        k = int(np.log2(U.shape[-1]))
        super().__init__(QubitUnitary(U, wires=wires[len(wires) - k:]), wires[: len(wires) - k],
                         control_values, id=id)

    @property
    def name(self):
        return "ControlledQubitUnitary"


class _Composite(Operator):
    is_hermitian = True

    def __init__(self, *operands, id=None):
        self.operands = tuple(operands)
        seen = []
        for o in self.operands:
            for w in o.wires:
                if w not in seen:
                    seen.append(w)
        self.wires = tuple(seen)
        self.data = tuple(d for o in self.operands for d in o.data)
        self.id = id
        self.hyperparameters = {}

    @property
    def num_params(self):
        return len(self.data)

    @property
    def batch_size(self):
        return None

    def __iter__(self):
        return iter(self.operands)

    def __len__(self):
        return len(self.operands)

    def __getitem__(self, i):
        return self.operands[i]

    @property
    def has_overlapping_wires(self):
        ws = [w for o in self.operands for w in o.wires]
        return len(ws) != len(set(ws))

    def map_wires(self, wire_map):
        return type(self)(*[o.map_wires(wire_map) for o in self.operands])


class Prod(_Composite):
    """op_math/prod.py:153."""

    def matrix(self, wire_order=None):
        wo = self.wires if wire_order is None else tuple(wire_order)
        m = np.eye(2 ** len(wo), dtype=complex)
        for o in self.operands:
            m = m @ np.asarray(o.matrix(wire_order=wo))
        return m

    @property
    def pauli_rep(self):
        ps = PauliSentence({PauliWord({}): 1.0})
        for o in self.operands:
            r = o.pauli_rep
            if r is None:
                return None
            ps = ps @ r
        return ps

    @property
    def has_diagonalizing_gates(self):
        return (not self.has_overlapping_wires) and all(
            o.has_diagonalizing_gates for o in self.operands)

    def diagonalizing_gates(self):
        if self.has_overlapping_wires:
            raise NotImplementedError("Prod with overlapping wires has no diagonalizing gates here")
        return [g for o in self.operands for g in o.diagonalizing_gates()]

    def eigvals(self):
        if self.has_overlapping_wires:
            return np.linalg.eigvalsh(self.matrix())
        ev = np.array([1.0])
        for o in self.operands:
            ev = np.kron(ev, np.asarray(o.eigvals()))
        return ev

    def terms(self):
        return [1.0], [self]


class SProd(Operator):
    """op_math/sprod.py:80."""
    is_hermitian = True

    def __init__(self, scalar, base, id=None):
        self.scalar = scalar
        self.base = base
        self.wires = base.wires
        self.data = (scalar,) + tuple(base.data)
        self.id = id
        self.hyperparameters = {}

    @property
    def num_params(self):
        return len(self.data)

    @property
    def batch_size(self):
        return None

    def matrix(self, wire_order=None):
        return self.scalar * np.asarray(self.base.matrix(wire_order=wire_order))

    @property
    def pauli_rep(self):
        r = self.base.pauli_rep
        return None if r is None else r.scale(self.scalar)

    @property
    def has_diagonalizing_gates(self):
        return self.base.has_diagonalizing_gates

    def diagonalizing_gates(self):
        return self.base.diagonalizing_gates()

    def eigvals(self):
        return self.scalar * np.asarray(self.base.eigvals())

    def terms(self):
        cs, os_ = self.base.terms() if hasattr(self.base, "terms") else ([1.0], [self.base])
        return [self.scalar * c for c in cs], os_

    def map_wires(self, wire_map):
        return SProd(self.scalar, self.base.map_wires(wire_map))


class Sum(_Composite):
    """op_math/sum.py:121."""

    def matrix(self, wire_order=None):
        wo = self.wires if wire_order is None else tuple(wire_order)
        return sum(np.asarray(o.matrix(wire_order=wo)) for o in self.operands)

    @property
    def pauli_rep(self):
        ps = PauliSentence()
        for o in self.operands:
            r = o.pauli_rep
            if r is None:
                return None
            ps = ps + r
        return ps

    def terms(self):
        cs, os_ = [], []
        for o in self.operands:
            if hasattr(o, "terms"):
                c2, o2 = o.terms()
                cs += list(c2); os_ += list(o2)
            else:
                cs.append(1.0); os_.append(o)
        return cs, os_

    def eigvals(self):
        return np.linalg.eigvalsh(self.matrix())


class LinearCombination(Sum):
    """op_math/linear_combination.py:33 — ``sum_i c_i O_i`` (``qml.Hamiltonian``)."""

    def __init__(self, coeffs, observables, id=None):
        if len(coeffs) != len(observables):
            raise ValueError("Could not create valid LinearCombination; number of coefficients "
                             "and operators does not match.")
        self._coeffs = list(coeffs)
        self._ops = list(observables)
        super().__init__(*[SProd(c, o) for c, o in zip(coeffs, observables)], id=id)

    def terms(self):
        return list(self._coeffs), list(self._ops)

    def map_wires(self, wire_map):
        return LinearCombination(self._coeffs, [o.map_wires(wire_map) for o in self._ops])


Hamiltonian = LinearCombination


def dot(coeffs, ops_):
    """``qml.dot`` for numeric coefficients and operators."""
    return LinearCombination(list(coeffs), list(ops_))


def matrix(op: Operator, wire_order=None) -> np.ndarray:
    """``qml.matrix(op, wire_order)``."""
    return np.asarray(op.matrix(wire_order=wire_order))


def generator_matrix(op: Operator) -> np.ndarray:
    """Matrix of ``op``'s generator on ``op.wires`` — the ``qml.matrix(qml.generator(op,
    format="observable"), wire_order=op.wires)`` of pennylane/operation.py:59."""
    return np.asarray(op.generator().matrix(wire_order=op.wires))


def operation_derivative(op: Operator) -> np.ndarray:
    """pennylane/operation.py:40-60: ``1j * G @ U``."""
    return 1j * generator_matrix(op) @ np.asarray(op.matrix())


def tensor_product_wires(*ops_):
    return tuple(itertools.chain.from_iterable(o.wires for o in ops_))


class Snapshot(Operator):
    """Debugger snapshot (pennylane/ops/meta.py:158-300): records ``measurement`` of the state at
    this point of the circuit into the device's debugger (apply_operation.py:883-917); a no-op
    without an active debugger."""

    num_params = 0
    has_matrix = False

    def __init__(self, tag=None, measurement=None, shots="workflow"):
        from .measurements import MeasurementProcess, StateMP
        from .tape import Shots

        if tag is not None and not isinstance(tag, (str, int)):
            raise ValueError("Snapshot tags can only be of type 'str'")
        if measurement is None:
            measurement = StateMP()
        if not isinstance(measurement, MeasurementProcess):
            raise ValueError(f"The measurement {measurement.__class__.__name__} is not supported "
                             f"as it is not an instance of {MeasurementProcess}")
        if isinstance(measurement, StateMP) and isinstance(shots, str) and shots == "workflow":
            shots = None                                   # always analytic with state
        super().__init__(wires=measurement.wires)
        self.hyperparameters = {
            "tag": tag, "measurement": measurement,
            "shots": shots if (isinstance(shots, str) and shots == "workflow") else Shots(shots)}

    @property
    def tag(self):
        return self.hyperparameters["tag"]

    def decomposition(self):
        return []

    def adjoint(self):
        return Snapshot(**self.hyperparameters)

    def map_wires(self, wire_map: dict):
        return Snapshot(tag=self.tag, measurement=self.hyperparameters["measurement"].map_wires(wire_map),
                        shots=self.hyperparameters["shots"])

    def __repr__(self):
        return (f"<Snapshot: tag={self.tag}, measurement={self.hyperparameters['measurement']}, "
                f"shots={self.hyperparameters['shots']}>")


__all__ = [n for n, v in list(globals().items())
           if isinstance(v, type) and issubclass(v, Operator)] + [
    "adjoint", "ctrl", "dot", "matrix", "generator_matrix", "operation_derivative",
    "expand_matrix", "pauli_word_op", "X", "Y", "Z", "H", "CPhase", "Hamiltonian"]
