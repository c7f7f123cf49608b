"""Mid-circuit measurements — mirror of the operator classes the device consumes.

Reference: pennylane/ops/mid_measure/mid_measure.py (``MidMeasure`` :100-190, ``measure``
:192), pennylane/ops/mid_measure/measurement_value.py (``MeasurementValue`` :35,
``concretize`` :201-204) and pennylane/ops/op_math/condition.py (``Conditional`` :113-190).
The engine applies them natively (``StateVector.apply_mid_measure`` →
``b200q_probs`` + ``b200q_collapse``; apply_operation.py:355-497).
"""
from __future__ import annotations

import itertools

import numpy as np

from .ops import Operator, _wires_tuple

_uid = itertools.count()


class MidMeasure(Operator):
    """Projective measurement of one wire in the computational basis (mid_measure.py:100).

    ``name`` is ``"MidMeasureMP"`` as in the reference (:136); instances hash by identity so
    they can key the ``mid_measurements`` dictionary (apply_operation.py:473)."""

    num_wires = 1
    num_params = 0
    has_matrix = False

    def __init__(self, wires=None, reset: bool = False, postselect=None, id=None, meas_uid=None):
        super().__init__(wires=wires, id=id)
        self.hyperparameters = {"reset": bool(reset), "postselect": postselect,
                                "meas_uid": meas_uid if meas_uid is not None else next(_uid)}

    @property
    def name(self) -> str:
        return "MidMeasureMP"

    @property
    def reset(self) -> bool:
        return self.hyperparameters["reset"]

    @property
    def postselect(self):
        return self.hyperparameters["postselect"]

    @property
    def meas_uid(self):
        return self.hyperparameters["meas_uid"]

    def map_wires(self, wire_map: dict):
        # the same measurement after relabelling: conditionals and terminal MCM samples refer to
        # it by identity, so standard-wire mapping must not break that link
        new = MidMeasure(tuple(wire_map.get(w, w) for w in self.wires), reset=self.reset,
                         postselect=self.postselect, id=self.id, meas_uid=self.meas_uid)
        return new

    def __hash__(self):
        return hash((type(self), self.meas_uid))

    def __eq__(self, other):
        return isinstance(other, MidMeasure) and other.meas_uid == self.meas_uid

    def __repr__(self):
        return f"MidMeasure(wires={list(self.wires)}, reset={self.reset}, postselect={self.postselect})"


class MeasurementValue:
    """Lazy classical value depending on mid-circuit measurements (measurement_value.py:35)."""

    name = "MeasurementValue"

    def __init__(self, measurements, processing_fn=None):
        self.measurements = list(measurements)
        self._processing_fn = processing_fn

    @property
    def processing_fn(self):
        return (lambda *x: x[0] if len(x) == 1 else x) if self._processing_fn is None \
            else self._processing_fn

    @property
    def wires(self):
        seen = []
        for m in self.measurements:
            for w in m.wires:
                if w not in seen:
                    seen.append(w)
        return tuple(seen)

    def concretize(self, measurements: dict):
        """measurement_value.py:201-204."""
        return self.processing_fn(*(measurements[m] for m in self.measurements))

    # ---- the arithmetic used to build conditions (measurement_value.py:95-190) ---------------
    def _apply(self, fn):
        pf = self.processing_fn
        return MeasurementValue(self.measurements, lambda *x: fn(pf(*x)))

    def _merge(self, other: "MeasurementValue"):
        merged = list(self.measurements)
        for m in other.measurements:
            if m not in merged:
                merged.append(m)
        merged.sort(key=lambda m: m.meas_uid)
        i1 = [merged.index(m) for m in self.measurements]
        i2 = [merged.index(m) for m in other.measurements]
        f1, f2 = self.processing_fn, other.processing_fn
        return MeasurementValue(
            merged, lambda *x: (f1(*(x[i] for i in i1)), f2(*(x[i] for i in i2))))

    def _binary(self, other, fn):
        if isinstance(other, MeasurementValue):
            return self._merge(other)._apply(lambda t: fn(t[0], t[1]))
        return self._apply(lambda v: fn(v, other))

    def __invert__(self):
        return self._apply(np.logical_not)

    def __eq__(self, other):
        return self._binary(other, lambda a, b: a == b)

    def __ne__(self, other):
        return self._binary(other, lambda a, b: a != b)

    def __add__(self, other):
        return self._binary(other, lambda a, b: a + b)

    __radd__ = __add__

    def __mul__(self, other):
        return self._binary(other, lambda a, b: a * b)

    __rmul__ = __mul__

    def __sub__(self, other):
        return self._binary(other, lambda a, b: a - b)

    def __and__(self, other):
        return self._binary(other, np.logical_and)

    def __or__(self, other):
        return self._binary(other, np.logical_or)

    def __lt__(self, other):
        return self._binary(other, lambda a, b: a < b)

    def __gt__(self, other):
        return self._binary(other, lambda a, b: a > b)

    def __hash__(self):
        return id(self)

    def __bool__(self):
        raise ValueError("The truth value of a MeasurementValue is undefined. To condition on a "
                         "MeasurementValue, please use cond instead.")

    def map_wires(self, wire_map: dict):
        return MeasurementValue([m.map_wires(wire_map) for m in self.measurements],
                                self._processing_fn)


class Conditional(Operator):
    """``then_op`` applied iff ``expr`` evaluates truthy on the sampled mid-circuit values
    (condition.py:113-190; apply_operation.py:355-411)."""

    has_matrix = False

    def __init__(self, expr: MeasurementValue, then_op: Operator, id=None):
        self.meas_val = expr
        self.base = then_op
        self.wires = _wires_tuple(then_op.wires)
        self.data = tuple(then_op.data)
        self.id = id
        self.hyperparameters = {}

    @property
    def name(self) -> str:
        return f"Conditional({self.base.name})"

    @property
    def num_params(self):
        return self.base.num_params

    @property
    def batch_size(self):
        return self.base.batch_size

    def map_wires(self, wire_map: dict):
        return Conditional(self.meas_val.map_wires(wire_map), self.base.map_wires(wire_map),
                           id=self.id)

    def adjoint(self):
        return Conditional(self.meas_val, self.base.adjoint())

    def __repr__(self):
        return f"Conditional({self.base!r})"


def measure(wires, reset: bool = False, postselect=None) -> MeasurementValue:
    """mid_measure.py:192: the value object; the operator is ``.measurements[0]``."""
    w = _wires_tuple(wires)
    if len(w) != 1:
        raise ValueError("Only a single qubit can be measured in the middle of the circuit")
    return MeasurementValue([MidMeasure(w, reset=reset, postselect=postselect)])


def cond(condition: MeasurementValue, then_op: Operator) -> Conditional:
    """``qml.cond`` for one already-built operator (condition.py:220 keeps a callable)."""
    return Conditional(condition, then_op)


def is_mcm(op) -> bool:
    return getattr(op, "name", "") == "MidMeasureMP"


def is_conditional(op) -> bool:
    return getattr(op, "name", "").startswith("Conditional") and hasattr(op, "meas_val")


__all__ = ["MidMeasure", "MeasurementValue", "Conditional", "measure", "cond", "is_mcm",
           "is_conditional"]
