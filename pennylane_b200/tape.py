"""``QuantumScript`` / ``Shots`` — mirror of the data model the device receives at its boundary.

Reference: pennylane/core/qscript.py:84 (``operations`` :242, ``measurements`` :283,
``observables`` :256, ``shots`` :384, ``batch_size`` :309, ``trainable_params`` :554,
``num_preps`` :394, ``hash`` :193, ``map_to_standard_wires`` :1015) and
pennylane/core/shots.py:48 (``total_shots``, ``shot_vector``, ``bins`` :288,
``has_partitioned_shots`` :271, ``num_copies``).
"""
from __future__ import annotations

import hashlib
from collections import namedtuple

import numpy as np

ShotCopies = namedtuple("ShotCopies", ["shots", "copies"])


class Shots:
    def __init__(self, shots=None):
        if isinstance(shots, Shots):
            self.total_shots, self.shot_vector = shots.total_shots, shots.shot_vector
            return
        if shots is None:
            self.total_shots, self.shot_vector = None, ()
            return
        if isinstance(shots, (int, np.integer)):
            if shots < 1:
                raise ValueError(f"Shots must be a positive integer; got {shots}")
            self.total_shots = int(shots)
            self.shot_vector = (ShotCopies(int(shots), 1),)
            return
        vec = []
        for s in shots:
            if isinstance(s, (tuple, list)):
                n, c = int(s[0]), int(s[1])
            else:
                n, c = int(s), 1
            if n < 1 or c < 1:
                raise ValueError(f"Shots must be positive integers; got {shots}")
            if vec and vec[-1].shots == n:
                vec[-1] = ShotCopies(n, vec[-1].copies + c)
            else:
                vec.append(ShotCopies(n, c))
        self.shot_vector = tuple(vec)
        self.total_shots = sum(s * c for s, c in vec)

    def __bool__(self):
        return self.total_shots is not None

    def __iter__(self):
        for s, c in self.shot_vector:
            for _ in range(c):
                yield s

    def __eq__(self, other):
        return isinstance(other, Shots) and self.shot_vector == other.shot_vector

    def __hash__(self):
        return hash(self.shot_vector)

    def __repr__(self):
        return f"Shots(total_shots={self.total_shots}, shot_vector={self.shot_vector})"

    @property
    def has_partitioned_shots(self):
        if not self:
            return False
        return len(self.shot_vector) > 1 or self.shot_vector[0].copies > 1

    @property
    def num_copies(self):
        return sum(c for _, c in self.shot_vector)

    def bins(self):
        lower = 0
        for s, c in self.shot_vector:
            for _ in range(c):
                yield lower, lower + s
                lower += s


class FlexShots(Shots):
    """``_FlexShots`` (devices/qubit/simulate.py:104-117): shots that may be zero — what is left of
    a shot budget after a postselecting projector (``_postselection_postprocess`` :159-167)."""

    def __init__(self, shots=None):
        if isinstance(shots, (int, np.integer)):
            vec = [ShotCopies(int(shots), 1)]
        else:
            vec = []
            for s in shots:
                n, c = (int(s[0]), int(s[1])) if isinstance(s, (tuple, list)) else (int(s), 1)
                if vec and vec[-1].shots == n:
                    vec[-1] = ShotCopies(n, vec[-1].copies + c)
                else:
                    vec.append(ShotCopies(n, c))
        self.shot_vector = tuple(vec)
        self.total_shots = sum(s * c for s, c in vec)


class QuantumScript:
    """An executable circuit: operations, measurements, shots, trainable parameter indices."""

    def __init__(self, ops=(), measurements=(), shots=None, trainable_params=None):
        self.operations = list(ops)
        self.measurements = list(measurements)
        self.shots = shots if isinstance(shots, Shots) else Shots(shots)
        self._trainable_params = None if trainable_params is None else list(trainable_params)

    # ---- structure -------------------------------------------------------------------------
    def __len__(self):
        return len(self.operations) + len(self.measurements)

    def __getitem__(self, i):
        return (self.operations + self.measurements)[i]

    @property
    def observables(self):
        return [m.obs if m.obs is not None else m for m in self.measurements]

    @property
    def op_wires(self):
        seen = []
        for op in self.operations:
            for w in op.wires:
                if w not in seen:
                    seen.append(w)
        return tuple(seen)

    @property
    def wires(self):
        seen = list(self.op_wires)
        for m in self.measurements:
            for w in m.wires:
                if w not in seen:
                    seen.append(w)
        return tuple(seen)

    @property
    def num_wires(self):
        return len(self.wires)

    @property
    def num_preps(self):
        n = 0
        for op in self.operations:
            if hasattr(op, "state_vector"):
                n += 1
            else:
                break
        return n

    @property
    def batch_size(self):
        bs = None
        for op in self.operations:
            b = op.batch_size
            if b is not None:
                if bs is not None and b != bs:
                    raise ValueError("The batch sizes of the quantum script operations do not "
                                     f"match, they include {bs} and {b}.")
                bs = b
        return bs

    # ---- parameters ------------------------------------------------------------------------
    def _par_info(self):
        info = []
        for i, op in enumerate(self.operations):
            for j in range(len(op.data)):
                info.append(("op", i, j))
        for i, m in enumerate(self.measurements):
            if m.obs is not None:
                for j in range(len(m.obs.data)):
                    info.append(("obs", i, j))
        return info

    @property
    def trainable_params(self):
        if self._trainable_params is None:
            return list(range(len(self._par_info())))
        return list(self._trainable_params)

    @trainable_params.setter
    def trainable_params(self, idx):
        n = len(self._par_info())
        if any(not isinstance(i, (int, np.integer)) or i < 0 or i >= n for i in idx):
            raise ValueError("Argument indices must be non-negative integers smaller than the "
                             f"number of parameters {n}.")
        self._trainable_params = sorted(set(int(i) for i in idx))

    def get_parameters(self, trainable_only=True, operations_only=False):
        params = []
        for idx, (kind, i, j) in enumerate(self._par_info()):
            if operations_only and kind != "op":
                continue
            if trainable_only and idx not in self.trainable_params:
                continue
            params.append(self.operations[i].data[j] if kind == "op"
                          else self.measurements[i].obs.data[j])
        return params

    @property
    def num_params(self):
        return len(self.trainable_params)

    # ---- transforms ------------------------------------------------------------------------
    def copy(self, **updates):
        return QuantumScript(
            updates.get("operations", updates.get("ops", self.operations)),
            updates.get("measurements", self.measurements),
            shots=updates.get("shots", self.shots),
            trainable_params=updates.get("trainable_params", self._trainable_params),
        )

    def map_to_standard_wires(self):
        """qscript.py:1015 — relabel wires to 0..n-1: operation wires (in order of appearance)
        first, measurement-only wires after."""
        wires = self.wires
        op_w = list(self.op_wires)
        meas_only_set = set(wires) - set(op_w)
        n_op = len(op_w)
        # qscript.py:1076-1082: op wires already 0..k-1 followed by measurement-only wires
        if set(op_w) == set(range(n_op)) and meas_only_set == set(range(n_op, n_op + len(meas_only_set))):
            return self
        meas_only = [w for w in wires if w not in op_w]
        wire_map = {w: i for i, w in enumerate(op_w + meas_only)}
        return QuantumScript([op.map_wires(wire_map) for op in self.operations],
                             [m.map_wires(wire_map) for m in self.measurements],
                             shots=self.shots, trainable_params=self._trainable_params)

    @property
    def hash(self):
        h = hashlib.sha1()
        for op in self.operations:
            h.update(op.name.encode())
            h.update(repr(op.wires).encode())
            for d in op.data:
                h.update(np.asarray(d).tobytes())
            h.update(repr(sorted((k, repr(v)) for k, v in op.hyperparameters.items()
                                 if k != "base")).encode())
        for m in self.measurements:
            h.update(repr(m).encode())
            if m.obs is not None:
                for d in m.obs.data:
                    h.update(np.asarray(d).tobytes())
        h.update(repr(self.shots).encode())
        return int(h.hexdigest()[:16], 16)

    def __repr__(self):
        return f"<QuantumScript: wires={list(self.wires)}, ops={len(self.operations)}, " \
               f"measurements={len(self.measurements)}>"


QuantumTape = QuantumScript
