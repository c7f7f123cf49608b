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

NOTE: PennyLane cannot be imported in the build container (autograd /
autoray / rustworkx are absent), so this file cannot be executed here.  What IS checked here,
statically against the reference sources (tests/test_plugin_conformance.py): every name this
module imports from ``pennylane`` exists at that path, every overridden ``Device`` method has the
reference's signature, and ``preprocess_transforms`` adds the transforms of
``DefaultQubit.preprocess_transforms`` in the same order with the same keywords.
"""
from __future__ import annotations

from dataclasses import replace

import numpy as np
import pennylane as qml
from pennylane import math
from pennylane.core.transforms import CompilePipeline
from pennylane.devices import Device, ExecutionConfig
from pennylane.devices.default_qubit import (ALL_DQ_GATES, ALL_DQ_GATES_PLUS_MCM,
                                             _add_adjoint_transforms, _conditional_broadcast_expand,
                                             _supports_adjoint, accepted_analytic_measurement,
                                             accepted_sample_measurement,
                                             allow_mcms_stopping_condition, no_counts,
                                             no_mcms_stopping_condition)
from pennylane.devices.modifiers import simulator_tracking, single_tape_support
from pennylane.devices.preprocess import (decompose, device_resolve_dynamic_wires, no_sampling,
                                          validate_device_wires, validate_measurements)
from pennylane.exceptions import DeviceError
from pennylane.transforms import broadcast_expand, defer_measurements, dynamic_one_shot

from . import adjoint as _adjoint
from . import simulate as _sim
from .device import stopping_condition as _engine_accepts

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
            raise DeviceError(f"Measurement {mp} is not supported on b200.qubit")
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
            raise DeviceError("b200.qubit does not support max_workers; run one device per GPU")
        super().__init__(wires=wires, shots=shots)
        seed = np.random.randint(0, high=10000000) if isinstance(seed, str) and seed == "global" else seed
        self._rng = np.random.default_rng(seed)          # default_qubit.py:562-570
        self._c_dtype = np.dtype(c_dtype)
        self._fusion = int(fusion)
        self._exact_sampling = bool(exact_sampling)
        self._debugger = None
        self._state_cache = None

    # ---- capability + configuration (default_qubit.py:572-608, 683-760) -----------------------
    def supports_derivatives(self, execution_config=None, circuit=None):
        """default_qubit.py:572-608 without the backprop branch (amplitudes live in CUDA kernels:
        nothing for an autodiff framework to trace)."""
        if execution_config is None:
            return True
        if execution_config.gradient_method in {"adjoint", "best"}:
            return _supports_adjoint(circuit, device_wires=self.wires, device_name=self.name)
        return False

    supports_jvp = supports_derivatives
    supports_vjp = supports_derivatives

    def setup_execution_config(self, config=None, circuit=None):
        """default_qubit.py:683-737; "best" resolves to adjoint (there is no backprop)."""
        config = config or ExecutionConfig()
        updated_values = {}
        for option in config.device_options:
            if option not in self._device_options:
                raise DeviceError(f"device option {option} not present on {self}")
        gradient_method = config.gradient_method
        if config.gradient_method == "best":
            gradient_method = "adjoint"
            updated_values["gradient_method"] = gradient_method
        if config.use_device_gradient is None:
            updated_values["use_device_gradient"] = gradient_method == "adjoint"
        if config.use_device_jacobian_product is None:
            updated_values["use_device_jacobian_product"] = gradient_method == "adjoint"
        if config.grad_on_execution is None:
            updated_values["grad_on_execution"] = gradient_method == "adjoint"
        updated_values["device_options"] = dict(config.device_options)  # copy
        for option in self._device_options:
            if option not in updated_values["device_options"]:
                updated_values["device_options"][option] = getattr(self, f"_{option}")
        updated_values["mcm_config"] = self._setup_mcm_config(config.mcm_config, circuit)
        return replace(config, **updated_values)

    def _setup_mcm_config(self, mcm_config, tape):
        """default_qubit.py:739-760 (all three methods are native on this device)."""
        final_mcm_method = mcm_config.mcm_method
        if mcm_config.mcm_method is None:
            final_mcm_method = "one-shot" if getattr(tape, "shots", None) else "deferred"
        elif mcm_config.mcm_method == "device":
            final_mcm_method = "tree-traversal"
        supported_methods = {"one-shot", "deferred", "tree-traversal"}
        if final_mcm_method not in supported_methods:
            raise DeviceError(f"mcm_method {final_mcm_method} not supported on b200.qubit. "
                              f"Supported methods are {supported_methods}")
        if mcm_config.postselect_mode == "fill-shots" and final_mcm_method != "deferred":
            raise DeviceError(
                "Using postselect_mode='fill-shots' is only supported with mcm_method='deferred'.")
        return replace(mcm_config, mcm_method=final_mcm_method)

    def preprocess_transforms(self, execution_config=None):
        """default_qubit.py:611-679, transform for transform (tests/test_plugin_conformance.py
        compares the two functions' ``add_transform`` sequences).  Left out: the
        ``validate_multiprocessing_workers`` step (``max_workers`` is rejected in ``__init__``).
        The stopping conditions are the reference's, narrowed by what the engine can apply
        (dense matrices up to 10 wires)."""
        config = execution_config or ExecutionConfig()

        compile_pipeline = CompilePipeline()
        target_gate_set = ALL_DQ_GATES

        if config.interface == math.Interface.JAX_JIT:
            compile_pipeline.add_transform(no_counts)

        if config.mcm_config.mcm_method == "deferred":
            compile_pipeline.add_transform(defer_measurements, allow_postselect=True)
            _reference_condition = no_mcms_stopping_condition
        else:
            _reference_condition = allow_mcms_stopping_condition
            target_gate_set = ALL_DQ_GATES_PLUS_MCM

        def _stopping_condition(op):
            return _reference_condition(op) and _engine_accepts(op)

        compile_pipeline.add_transform(
            decompose,
            stopping_condition=_stopping_condition,
            device_wires=self.wires,
            target_gates=target_gate_set,
            name=self.name,
        )
        _allow_resets = config.mcm_config.mcm_method != "deferred"
        compile_pipeline.add_transform(
            device_resolve_dynamic_wires, wires=self.wires, allow_resets=_allow_resets
        )
        compile_pipeline.add_transform(validate_device_wires, self.wires, name=self.name)
        compile_pipeline.add_transform(
            validate_measurements,
            analytic_measurements=accepted_analytic_measurement,
            sample_measurements=accepted_sample_measurement,
            name=self.name,
        )
        compile_pipeline.add_transform(_conditional_broadcast_expand)
        if config.mcm_config.mcm_method == "tree-traversal":
            compile_pipeline.add_transform(broadcast_expand)

        if config.mcm_config.mcm_method == "one-shot":
            compile_pipeline.add_transform(
                dynamic_one_shot, postselect_mode=config.mcm_config.postselect_mode
            )

        if config.gradient_method == "backprop":
            compile_pipeline.add_transform(no_sampling, name="backprop + b200.qubit")

        if config.gradient_method == "adjoint":
            # the reference's own helper: no_sampling, decompose(adjoint_ops), validate_observables,
            # validate_measurements, adjoint_state_measurements, broadcast_expand,
            # validate_adjoint_trainable_params (default_qubit.py:315-349)
            _add_adjoint_transforms(
                compile_pipeline,
                device_vjp=config.use_device_jacobian_product,
                device_wires=self.wires,
                target_gates=target_gate_set,
            )
        return compile_pipeline

    # ---- execution (default_qubit.py:763-1071) --------------------------------------------------
    def _opts(self, config):
        o = (config.device_options if config else {}) or {}
        return (o.get("rng", self._rng), o.get("c_dtype", self._c_dtype),
                o.get("fusion", self._fusion), o.get("exact_sampling", self._exact_sampling))

    def execute(self, circuits, execution_config=None):
        """default_qubit.py:763-846.  With ``use_device_jacobian_product`` the final state of every
        circuit is kept (``_state_cache``, :772) for the ``compute_vjp`` that follows (:1021-1029)."""
        if execution_config is None:
            execution_config = ExecutionConfig()
        rng, dt, fusion, exact = self._opts(execution_config)
        self._state_cache = {} if execution_config.use_device_jacobian_product else None
        mcm_method = execution_config.mcm_config.mcm_method
        return tuple(_sim.simulate(_Tape(c), rng=rng, dtype=dt, exact_sampling=exact, fusion=fusion,
                                   debugger=self._debugger, state_cache=self._state_cache,
                                   mcm_method=mcm_method)
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
        """default_qubit.py:1000-1034: the reverse sweep starts from the state the preceding
        ``execute`` left in ``_state_cache`` (keyed by the circuit's hash) when there is one."""
        _, dt, fusion, _ = self._opts(execution_config)
        cache = getattr(self, "_state_cache", None) or {}
        return tuple(_adjoint.adjoint_vjp(_Tape(c), t, dtype=dt, fusion=fusion, state=cache.get(c.map_to_standard_wires().hash))
                     for c, t in zip(circuits, cotangents))

    def execute_and_compute_vjp(self, circuits, cotangents, execution_config=None):
        return self.execute(circuits, execution_config), self.compute_vjp(circuits, cotangents, execution_config)
