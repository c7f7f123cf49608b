"""``qml.device("b200.qubit")`` — the genuine ``pennylane.devices.Device`` subclass.

Importing this module needs PennyLane (>= 0.44).  It is a thin adapter: PennyLane's own
preprocessing transforms (``pennylane/devices/preprocess.py``) build the pipeline exactly as
``DefaultQubit.preprocess_transforms`` does (``default_qubit.py:611-679``), and execution hands the
preprocessed ``QuantumScript`` objects — unchanged — to the same engine that the mirror classes in
this package drive (``simulate.py`` / ``adjoint.py``): the engine only duck-types the attributes
PennyLane's ``Operator`` / ``MeasurementProcess`` / ``QuantumScript`` already have (``name``,
``wires``, ``data``, ``matrix()``, ``hyperparameters``, ``generator()``, ``pauli_rep``,
``operations``, ``measurements``, ``shots``, ``trainable_params`` ...).

Entry point (``pyproject.toml`` of this package, group ``pennylane.plugins``,
``devices/device_constructor.py:30-57``)::

    [project.entry-points."pennylane.plugins"]
    "b200.qubit" = "pennylane_b200.pl_plugin:B200QubitDevice"

NOTE: PennyLane cannot be imported in the build container, so this file is
exercised only where a real install exists; the engine underneath is what the test-suite covers.
"""
from __future__ import annotations

import numpy as np
import pennylane as qml
from pennylane.devices import DefaultQubit, Device, ExecutionConfig
from pennylane.devices.modifiers import simulator_tracking, single_tape_support
from pennylane.devices.preprocess import (decompose, no_sampling, validate_adjoint_trainable_params,
                                          validate_device_wires, validate_measurements,
                                          validate_observables)
from pennylane.transforms.core import TransformProgram

from . import adjoint as _adjoint
from . import simulate as _sim
from .device import adjoint_observables, adjoint_ops, stopping_condition

_KIND = {"ExpectationMP": "expval", "VarianceMP": "var", "ProbabilityMP": "probs",
         "SampleMP": "sample", "CountsMP": "counts", "StateMP": "state",
         "DensityMatrixMP": "density_matrix", "PurityMP": "purity", "VnEntropyMP": "vn_entropy",
         "MutualInfoMP": "mutual_info", "ClassicalShadowMP": "shadow",
         "ShadowExpvalMP": "shadow_expval"}


class _MP:
    """Adapter giving a PennyLane measurement process the ``kind`` tag and the two-argument
    ``process_samples`` the engine uses."""

    def __init__(self, mp):
        self._mp = mp
        self.kind = _KIND.get(type(mp).__name__)
        if self.kind is None:
            raise qml.DeviceError(f"Measurement {mp} is not supported on b200.qubit")
        self.obs = mp.obs
        self.mv = getattr(mp, "mv", None)     # sampled mid-circuit value of a one-shot tape
        self.log_base = getattr(mp, "log_base", None)
        self.seed, self.H, self.k = (getattr(mp, a, None) for a in ("seed", "H", "k"))
        self._wires = getattr(mp, "_wires", None)      # MutualInfoMP: the two subsystems
        self.wires = tuple(mp.wires)

    def diagonalizing_gates(self):
        return self._mp.diagonalizing_gates()

    def eigvals(self):
        return self._mp.eigvals()

    def process_samples(self, samples, wire_order):
        return self._mp.process_samples(samples, qml.wires.Wires(list(wire_order)))


_sim.MEASUREMENT_ADAPTER = _MP


class _Tape:
    """View of a ``QuantumScript`` whose measurements are wrapped in :class:`_MP`."""

    def __init__(self, tape):
        self._t = tape
        self.measurements = [_MP(m) for m in tape.measurements]

    def __getattr__(self, name):
        return getattr(self._t, name)

    def map_to_standard_wires(self):
        return _Tape(self._t.map_to_standard_wires())


@simulator_tracking
@single_tape_support
class B200QubitDevice(Device):
    """Statevector simulator on one NVIDIA B200 behind PennyLane's device API.

    Keyword arguments mirror ``DefaultQubit`` (``wires``, ``shots``, ``seed``) plus ``c_dtype``,
    ``fusion`` and ``exact_sampling`` (see :class:`pennylane_b200.device.B200Qubit`).
    ``max_workers`` is rejected: the device owns a CUDA context (``default_qubit.py:810-829``
    forks a process pool).
    """

    pennylane_requires = ">=0.44"
    version = "0.1.0"
    author = "b200-qubit"
    _device_options = ("rng", "c_dtype", "fusion", "exact_sampling")

    @property
    def name(self):
        return "b200.qubit"

    def __init__(self, wires=None, shots=None, seed="global", c_dtype=np.complex128,
                 fusion: int = 1, exact_sampling: bool = True, max_workers=None):
        if max_workers is not None:
            raise qml.DeviceError("b200.qubit does not support max_workers; run one device per GPU")
        super().__init__(wires=wires, shots=shots)
        seed = np.random.randint(0, high=10000000) if isinstance(seed, str) and seed == "global" else seed
        self._rng = np.random.default_rng(seed)          # default_qubit.py:562-570
        self._c_dtype = np.dtype(c_dtype)
        self._fusion = int(fusion)
        self._exact = bool(exact_sampling)
        self._debugger = None

    # ---- capability + configuration (default_qubit.py:574-608, 683-733) -----------------------
    def supports_derivatives(self, execution_config=None, circuit=None):
        if execution_config is None:
            return True
        if execution_config.gradient_method not in ("adjoint", "best"):
            return False                                # no backprop: amplitudes live in CUDA kernels
        if circuit is None:
            return True
        return DefaultQubit().supports_derivatives(
            ExecutionConfig(gradient_method="adjoint"), circuit)

    supports_jvp = supports_derivatives
    supports_vjp = supports_derivatives

    def setup_execution_config(self, config=None, circuit=None):
        from dataclasses import replace
        config = config or ExecutionConfig()
        for option in config.device_options:
            if option not in self._device_options:
                raise qml.DeviceError(f"device option {option} not present on {self}")
        updated = {}
        method = "adjoint" if config.gradient_method == "best" else config.gradient_method
        updated["gradient_method"] = method
        if config.use_device_gradient is None:
            updated["use_device_gradient"] = method == "adjoint"
        if config.use_device_jacobian_product is None:
            updated["use_device_jacobian_product"] = method == "adjoint"
        if config.grad_on_execution is None:
            updated["grad_on_execution"] = method == "adjoint"
        opts = dict(config.device_options)
        opts.setdefault("rng", self._rng)
        opts.setdefault("c_dtype", self._c_dtype)
        opts.setdefault("fusion", self._fusion)
        opts.setdefault("exact_sampling", self._exact)
        updated["device_options"] = opts
        # _setup_mcm_config, default_qubit.py:740-760: one-shot with shots, deferred without;
        # tree-traversal is not built on this device
        mcm = config.mcm_config
        method = mcm.mcm_method
        if method is None:
            method = "one-shot" if getattr(circuit, "shots", None) else "deferred"
        if method not in ("deferred", "one-shot"):
            raise qml.DeviceError(f"mcm_method {method} not supported on b200.qubit. "
                                  "Supported methods are 'deferred' and 'one-shot'.")
        if mcm.postselect_mode == "fill-shots" and method != "deferred":
            raise qml.DeviceError(
                "Using postselect_mode='fill-shots' is only supported with mcm_method='deferred'.")
        updated["mcm_config"] = replace(mcm, mcm_method=method)
        return replace(config, **updated)

    def preprocess_transforms(self, execution_config=None):
        config = execution_config or ExecutionConfig()
        prog = TransformProgram()
        one_shot = config.mcm_config.mcm_method == "one-shot"      # default_qubit.py:632-664
        if one_shot:
            accept = stopping_condition                 # MidMeasure / Conditional applied natively
        else:
            prog.add_transform(qml.defer_measurements, allow_postselect=False)

            def accept(op):
                return op.name != "MidMeasureMP" and stopping_condition(op)
        prog.add_transform(validate_device_wires, self.wires, name=self.name)
        prog.add_transform(decompose, stopping_condition=accept, name=self.name)
        prog.add_transform(validate_measurements, name=self.name)
        prog.add_transform(validate_observables, lambda o: True, name=self.name)
        if one_shot:
            prog.add_transform(qml.transforms.dynamic_one_shot,
                               postselect_mode=config.mcm_config.postselect_mode)
        if config.gradient_method == "adjoint":         # _add_adjoint_transforms :315-349
            name = "adjoint + b200.qubit"
            prog.add_transform(no_sampling, name=name)
            prog.add_transform(decompose, stopping_condition=adjoint_ops, name=name,
                               skip_initial_state_prep=False)
            prog.add_transform(validate_observables, adjoint_observables, name=name)
            prog.add_transform(qml.transforms.broadcast_expand)
            prog.add_transform(validate_adjoint_trainable_params)
        return prog

    # ---- execution (default_qubit.py:763-1071) --------------------------------------------------
    def _opts(self, config):
        o = (config.device_options if config else {}) or {}
        return (o.get("rng", self._rng), o.get("c_dtype", self._c_dtype),
                o.get("fusion", self._fusion), o.get("exact_sampling", self._exact))

    def execute(self, circuits, execution_config=None):
        rng, dt, fusion, exact = self._opts(execution_config)
        return tuple(_sim.simulate(_Tape(c), rng=rng, dtype=dt, exact_sampling=exact, fusion=fusion,
                                   debugger=self._debugger)
                     for c in circuits)

    def compute_derivatives(self, circuits, execution_config=None):
        _, dt, fusion, _ = self._opts(execution_config)
        return tuple(_adjoint.adjoint_jacobian(_Tape(c), dtype=dt, fusion=fusion) for c in circuits)

    def execute_and_compute_derivatives(self, circuits, execution_config=None):
        _, dt, fusion, _ = self._opts(execution_config)
        res, jacs = [], []
        for c in circuits:
            t = _Tape(c).map_to_standard_wires()
            jac, final = _adjoint.adjoint_jacobian(t, dtype=dt, return_state=True, fusion=fusion)
            res.append(_sim.measure_final_state(t, final, False))
            jacs.append(jac)
        return tuple(res), tuple(jacs)

    def compute_jvp(self, circuits, tangents, execution_config=None):
        _, dt, fusion, _ = self._opts(execution_config)
        return tuple(_adjoint.adjoint_jvp(_Tape(c), t, dtype=dt, fusion=fusion)
                     for c, t in zip(circuits, tangents))

    def execute_and_compute_jvp(self, circuits, tangents, execution_config=None):
        return self.execute(circuits, execution_config), self.compute_jvp(circuits, tangents, execution_config)

    def compute_vjp(self, circuits, cotangents, execution_config=None):
        _, dt, fusion, _ = self._opts(execution_config)
        return tuple(_adjoint.adjoint_vjp(_Tape(c), t, dtype=dt, fusion=fusion)
                     for c, t in zip(circuits, cotangents))

    def execute_and_compute_vjp(self, circuits, cotangents, execution_config=None):
        return self.execute(circuits, execution_config), self.compute_vjp(circuits, cotangents, execution_config)
