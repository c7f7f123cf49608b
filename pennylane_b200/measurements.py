"""Measurement processes — mirror of the part of ``pennylane.measurements`` the device consumes.

Reference: pennylane/measurements/{expval,var,probs,sample,counts,state}.py and
pennylane/core/measurements (``MeasurementProcess``: ``obs``, ``wires``, ``eigvals()``,
``diagonalizing_gates()``, ``process_samples``).  ``kind`` is the duck-typing tag used by the
engine and the oracle instead of ``isinstance`` checks.
"""
from __future__ import annotations

import numpy as np


class MeasurementProcess:
    kind = None

    def __init__(self, obs=None, wires=None):
        if obs is not None and wires is not None:
            raise ValueError("Cannot set the wires if an observable is provided.")
        # a MeasurementValue (mid-circuit measurement outcome) is carried as ``mv``, never as
        # the observable (pennylane/core/measurements: ``MeasurementProcess.mv``)
        self.mv = None
        if getattr(obs, "name", None) == "MeasurementValue":
            self.mv, obs = obs, None
            wires = self.mv.wires
        self.obs = obs
        if obs is not None:
            self.wires = tuple(obs.wires)
        elif wires is None:
            self.wires = ()
        elif isinstance(wires, (str, bytes)) or not hasattr(wires, "__iter__"):
            self.wires = (wires,)
        else:
            self.wires = tuple(wires)

    def __repr__(self):
        inner = repr(self.obs) if self.obs is not None else (
            f"mcm wires={list(self.wires)}" if self.mv is not None else f"wires={list(self.wires)}")
        return f"{self.kind}({inner})"

    def map_wires(self, wire_map):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        if self.mv is not None:
            new.mv = self.mv.map_wires(wire_map)
        if self.obs is not None:
            new.obs = self.obs.map_wires(wire_map)
            new.wires = tuple(new.obs.wires)
        else:
            new.wires = tuple(wire_map.get(w, w) for w in self.wires)
        return new

    def diagonalizing_gates(self):
        return [] if self.obs is None else list(self.obs.diagonalizing_gates())

    def eigvals(self):
        return None if self.obs is None else np.asarray(self.obs.eigvals())

    # ---- measurements/*.py process_samples -------------------------------------------------
    def _indices(self, samples, wire_order):
        wire_order = list(wire_order)
        wires = list(self.wires) if len(self.wires) else wire_order
        cols = [wire_order.index(w) for w in wires]
        sub = samples[..., cols]
        powers = 2 ** np.arange(len(wires))[::-1]
        return sub, sub @ powers, wires

    def process_samples(self, samples, wire_order):  # pragma: no cover - abstract
        raise NotImplementedError


class ExpectationMP(MeasurementProcess):
    """measurements/expval.py (process_samples :60-79, process_state :81-93)."""
    kind = "expval"

    def process_samples(self, samples, wire_order):
        _, idx, _ = self._indices(samples, wire_order)
        vals = np.asarray(self.eigvals())[idx]
        return np.squeeze(np.mean(vals, axis=-1))


class VarianceMP(MeasurementProcess):
    """measurements/var.py."""
    kind = "var"

    def process_samples(self, samples, wire_order):
        _, idx, _ = self._indices(samples, wire_order)
        vals = np.asarray(self.eigvals())[idx]
        return np.squeeze(np.var(vals, axis=-1))


class ProbabilityMP(MeasurementProcess):
    """measurements/probs.py (process_samples :70-99, process_state :101-135)."""
    kind = "probs"

    def process_samples(self, samples, wire_order):
        _, idx, wires = self._indices(samples, wire_order)
        dim = 2 ** len(wires)
        if idx.ndim == 1:
            return np.bincount(idx, minlength=dim) / idx.shape[0]
        return np.stack([np.bincount(i, minlength=dim) / i.shape[0] for i in idx])


class SampleMP(MeasurementProcess):
    """measurements/sample.py."""
    kind = "sample"

    def process_samples(self, samples, wire_order):
        sub, idx, _ = self._indices(samples, wire_order)
        if self.obs is None:
            return sub
        return np.asarray(self.eigvals())[idx]


class CountsMP(MeasurementProcess):
    """measurements/counts.py."""
    kind = "counts"

    def __init__(self, obs=None, wires=None, all_outcomes=False):
        super().__init__(obs, wires)
        self.all_outcomes = all_outcomes

    def process_samples(self, samples, wire_order):
        sub, idx, wires = self._indices(samples, wire_order)
        out = {}
        if self.obs is None:
            if self.all_outcomes:
                for i in range(2 ** len(wires)):
                    out[format(i, f"0{len(wires)}b")] = 0
            vals, cnts = np.unique(idx, return_counts=True)
            for v, c in zip(vals, cnts):
                out[format(int(v), f"0{len(wires)}b")] = int(c)
            return out
        ev = np.asarray(self.eigvals())
        if self.all_outcomes:
            for e in np.unique(ev):
                out[float(e)] = 0
        vals, cnts = np.unique(ev[idx], return_counts=True)
        for v, c in zip(vals, cnts):
            out[float(v)] = int(c)
        return out


class StateMP(MeasurementProcess):
    """measurements/state.py."""
    kind = "state"


class DensityMatrixMP(MeasurementProcess):
    """``qml.density_matrix(wires)`` (measurements/state.py ``DensityMatrixMP``): the reduced
    density matrix over ``wires`` via ``reduce_statevector`` (math/quantum.py:386-487)."""
    kind = "density_matrix"


class PurityMP(MeasurementProcess):
    """measurements/purity.py:50-54: tr(rho_wires^2)."""
    kind = "purity"


class VnEntropyMP(MeasurementProcess):
    """measurements/vn_entropy.py:65-67: -tr(rho log rho) / log(base)."""
    kind = "vn_entropy"

    def __init__(self, wires=None, log_base=None):
        super().__init__(wires=wires)
        self.log_base = log_base


class MutualInfoMP(MeasurementProcess):
    """measurements/mutual_info.py:92-100: S(A) + S(B) - S(AB)."""
    kind = "mutual_info"

    def __init__(self, wires0, wires1, log_base=None):
        w0 = tuple(wires0) if hasattr(wires0, "__iter__") and not isinstance(wires0, str) else (wires0,)
        w1 = tuple(wires1) if hasattr(wires1, "__iter__") and not isinstance(wires1, str) else (wires1,)
        if set(w0) & set(w1):
            raise ValueError("Subsystems for computing mutual information must not overlap.")
        super().__init__(wires=w0 + w1)
        self._wires = (w0, w1)
        self.log_base = log_base

    def map_wires(self, wire_map):
        new = super().map_wires(wire_map)
        new._wires = tuple(tuple(wire_map.get(w, w) for w in ws) for ws in self._wires)
        return new


class ClassicalShadowMP(MeasurementProcess):
    """measurements/classical_shadow.py: random single-qubit Pauli measurements, one basis
    choice per shot and wire; ``seed`` fixes the recipes (``np.random.RandomState(seed)``, :171)."""
    kind = "shadow"

    def __init__(self, wires, seed=None):
        super().__init__(wires=wires)
        self.seed = seed


class ShadowExpvalMP(MeasurementProcess):
    """measurements/classical_shadow.py:400-514: expectation value(s) of ``H`` estimated from a
    classical shadow taken on the wires of ``H``; ``k`` = median-of-means batches."""
    kind = "shadow_expval"

    def __init__(self, H, seed=None, k=1):
        hs = list(H) if isinstance(H, (list, tuple)) else [H]
        wires = []
        for h in hs:
            wires += [w for w in h.wires if w not in wires]
        super().__init__(wires=wires)
        self.H, self.seed, self.k = H, seed, k

    def map_wires(self, wire_map):
        new = super().map_wires(wire_map)
        new.H = [h.map_wires(wire_map) for h in self.H] if isinstance(self.H, (list, tuple)) \
            else self.H.map_wires(wire_map)
        return new


def shadow_expval(H, k=1, seed=None):
    return ShadowExpvalMP(H, seed=seed, k=k)


def classical_shadow(wires, seed=None):
    return ClassicalShadowMP(wires, seed=seed)


def density_matrix(wires):
    return DensityMatrixMP(wires=wires)


def purity(wires):
    return PurityMP(wires=wires)


def vn_entropy(wires, log_base=None):
    return VnEntropyMP(wires=wires, log_base=log_base)


def mutual_info(wires0, wires1, log_base=None):
    return MutualInfoMP(wires0, wires1, log_base=log_base)


def expval(op):
    return ExpectationMP(obs=op)


def var(op):
    return VarianceMP(obs=op)


def probs(wires=None, op=None):
    return ProbabilityMP(obs=op, wires=wires)


def sample(op=None, wires=None):
    return SampleMP(obs=op, wires=wires)


def counts(op=None, wires=None, all_outcomes=False):
    return CountsMP(obs=op, wires=wires, all_outcomes=all_outcomes)


def state():
    return StateMP()
