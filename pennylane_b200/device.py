"""``B200Qubit`` — the drop-in device, mirroring ``DefaultQubit``'s implementation of the
``qml.devices.Device`` boundary (pennylane/devices/default_qubit.py:352-1071,
pennylane/devices/device_api.py:57) method for method:

    execute                          default_qubit.py:763-829
    supports_derivatives             :574-608      (adjoint only, like lightning;
                                                    tests/devices/test_lightning_qubit.py:48-59)
    setup_execution_config           :683-733
    preprocess / preprocess_transforms :611-679 + _add_adjoint_transforms :315-349
    compute_derivatives              :832-852
    execute_and_compute_derivatives  :855-877
    compute_jvp / execute_and_compute_jvp   :903-960
    compute_vjp / execute_and_compute_vjp   :974-1071
    tracker bookkeeping              devices/modifiers/simulator_tracking.py:25-60

This class works on the repo's mirror of the PennyLane data model (``pennylane_b200.tape``,
``.ops``, ``.measurements``); ``pennylane_b200/pl_plugin.py`` wraps it in a genuine
``pennylane.devices.Device`` subclass when PennyLane is importable.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from . import adjoint as _adjoint
from . import simulate as _sim
from .one_shot import dynamic_one_shot
from .tape import QuantumScript


class DeviceError(Exception):
    """pennylane.exceptions.DeviceError."""


class QuantumFunctionError(Exception):
    """pennylane.exceptions.QuantumFunctionError."""


@dataclass(frozen=True)
class ExecutionConfig:
    """pennylane/devices/execution_config.py:186-282 (fields the device reads)."""
    grad_on_execution: bool | None = None
    use_device_gradient: bool | None = None
    use_device_jacobian_product: bool | None = None
    gradient_method: str | None = None
    gradient_keyword_arguments: dict = field(default_factory=dict)
    device_options: dict = field(default_factory=dict)
    interface: str | None = None
    derivative_order: int = 1
    convert_to_numpy: bool = True
    mcm_config: "MCMConfig" = None

    def __post_init__(self):
        if self.mcm_config is None:
            object.__setattr__(self, "mcm_config", MCMConfig())


@dataclass(frozen=True)
class MCMConfig:
    """pennylane/devices/execution_config.py:31-71."""
    mcm_method: str | None = None
    postselect_mode: str | None = None


class Tracker:
    """pennylane/devices/tracker.py:21 (``update`` :191, ``record`` :243)."""

    def __init__(self, callback=None):
        self.active = False
        self.callback = callback
        self.reset()

    def __enter__(self):
        self.active = True
        return self

    def __exit__(self, *exc):
        self.active = False

    def reset(self):
        self.totals, self.history, self.latest = {}, {}, {}

    def update(self, **kwargs):
        self.latest = kwargs
        for k, v in kwargs.items():
            self.history.setdefault(k, []).append(v)
            if isinstance(v, (int, float)) and v is not None:
                self.totals[k] = self.totals.get(k, 0) + v

    def record(self):
        if self.callback is not None:
            self.callback(totals=self.totals, history=self.history, latest=self.latest)


# gates applied natively by a kernel (everything with a matrix of <= 10 wires is; the names are
# the default.qubit gate set, default_qubit.py:79-125, plus what the engine special-cases)
_MAX_MATRIX_WIRES = 10


def stopping_condition(op) -> bool:
    """default_qubit.py:139-160: can the engine apply ``op`` directly?"""
    if hasattr(op, "state_vector"):
        return True
    if op.name in ("Snapshot", "Barrier", "Identity", "GlobalPhase", "MultiControlledX", "MultiRZ",
                   "PauliRot"):
        return True
    if op.name == "MidMeasureMP":
        return True                       # default_qubit.py:145-146, allow_mcms (one-shot method)
    if op.name.startswith("Conditional") and hasattr(op, "meas_val"):
        return stopping_condition(op.base)
    if op.name == "GroverOperator":
        return len(op.wires) < 9          # apply_operation.py:845: matrix below nine wires
    base = getattr(op, "base", None)
    if op.name.startswith("C(") and base is not None:
        # default_qubit.py:118 ("C(gate)" for every base gate): controls are masks for the
        # kernels, only the base block is dense
        return (bool(getattr(base, "has_matrix", False)) and len(base.wires) <= 2
                and len(op.control_wires) <= 16 and getattr(base, "batch_size", None) is None) \
            or (bool(getattr(op, "has_matrix", False)) and len(op.wires) <= _MAX_MATRIX_WIRES)
    return bool(getattr(op, "has_matrix", False)) and len(op.wires) <= _MAX_MATRIX_WIRES


def adjoint_ops(op, trainable: bool = True) -> bool:
    """default_qubit.py:286-292."""
    if op.name == "MidMeasureMP" or op.name.startswith("Conditional"):
        return False                      # default_qubit.py:288
    npar = len(op.data)
    return npar == 0 or not trainable or (npar == 1 and getattr(op, "has_generator", False))


def adjoint_observables(obs) -> bool:
    """default_qubit.py:295-297."""
    return bool(getattr(obs, "has_matrix", True))


def _decompose(tape: QuantumScript, accept, name: str, max_depth: int = 10) -> QuantumScript:
    """pennylane/devices/preprocess.py:269 (``decompose``) for the mirror operator classes.
    Trainable-parameter indices are re-derived: a decomposed op's parameters are trainable iff
    any parameter of the original op was."""
    old_train = set(tape.trainable_params)
    new_ops, new_train = [], []
    pidx = 0
    nidx = 0

    def expand(op, trainable, depth):
        nonlocal nidx
        if accept(op, trainable):
            new_ops.append(op)
            for _ in op.data:
                if trainable:
                    new_train.append(nidx)
                nidx += 1
            return
        if depth >= max_depth or not getattr(op, "has_decomposition", False):
            raise DeviceError(
                f"Operator {op} not supported with {name} and does not provide a decomposition.")
        for sub in op.decomposition():
            expand(sub, trainable, depth + 1)

    for op in tape.operations:
        npar = len(op.data)
        trainable = any((pidx + j) in old_train for j in range(npar))
        pidx += npar
        expand(op, trainable, 0)
    # observable parameters keep their (shifted) positions
    n_old_op_params = pidx
    for t in sorted(old_train):
        if t >= n_old_op_params:
            new_train.append(nidx + (t - n_old_op_params))
    return QuantumScript(new_ops, tape.measurements, shots=tape.shots, trainable_params=new_train)


class Debugger:
    """pennylane/debugging/snapshot.py:35-58 (``_SnapshotDebugger``): while the context is
    active the device records every ``Snapshot`` it meets in ``snapshots``."""

    def __init__(self, dev=None):
        self.snapshots = {}
        self.active = dev is None          # a free-standing debugger (tests) is active at once
        self.device = dev
        if dev is not None:
            dev._debugger = self

    def __enter__(self):
        self.active = True
        return self

    def __exit__(self, *exc):
        self.active = False
        if self.device is not None:
            self.device._debugger = None


class _PreprocessedBatch(tuple):
    """The tapes ``preprocess`` returns.  The reference returns a transform program whose
    application yields ``(tapes, postprocessing)`` (device_api.py:269-339); this mirror returns the
    tapes directly, so the one post-processing step the device pipeline owns — combining the
    per-shot results of a one-shot mid-circuit-measurement tape (dynamic_one_shot.py:172-181) —
    rides along as ``.postprocessing(results)``."""

    def __new__(cls, tapes, posts):
        self = super().__new__(cls, tapes)
        self._posts = tuple(posts)
        return self

    def postprocessing(self, results):
        return tuple(r if p is None else p(r) for r, p in zip(results, self._posts))


class B200Qubit:
    """Statevector simulator on one NVIDIA B200, API-compatible with ``default.qubit``.

    Args:
        wires (int, Iterable, None): device wires; ``None`` infers them per circuit.
        shots (int, Sequence, None): default shots (tapes carry their own, device_api.py:242).
        seed ("global", int, ...): seed of the NumPy ``Generator`` threaded through all
            executions (default_qubit.py:562-570).
        c_dtype: ``np.complex128`` (default) or ``np.complex64``.
        exact_sampling (bool): build the CDF in numpy's summation order (bit-exact shots) or
            with the parallel scan.
        fusion (int): 0 = one kernel per gate (rounding-stable parity mode); 1 = host fusion
            pass with single-qubit block merging + tile kernel (default); 2 = also two-qubit
            dense blocks.
        device: torch CUDA device.
    """

    name = "b200.qubit"
    short_name = "b200.qubit"
    pennylane_requires = ">=0.44"
    version = "0.1.0"
    author = "b200-qubit"
    _device_options = ("rng", "c_dtype", "exact_sampling", "fusion", "return_torch")

    def __init__(self, wires=None, shots=None, seed="global", c_dtype=np.complex128,
                 exact_sampling: bool = True, device=None, max_workers=None, fusion: int = 1,
                 return_torch: bool = False):
        if max_workers is not None:
            raise DeviceError("b200.qubit owns a CUDA context and does not support max_workers "
                              "(process pools); run one device per GPU instead.")
        if wires is None:
            self.wires = None
        elif isinstance(wires, int):
            self.wires = tuple(range(wires))
        else:
            self.wires = tuple(wires)
        from .tape import Shots
        self.shots = Shots(shots)
        seed = np.random.randint(0, high=10000000) if isinstance(seed, str) and seed == "global" else seed
        self._rng = np.random.default_rng(seed)
        self._c_dtype = np.dtype(c_dtype)
        self._exact_sampling = bool(exact_sampling)
        self._fusion = int(fusion)
        self._return_torch = bool(return_torch)     # state / probs / Jacobians stay on the device
        self._torch_device = device
        self._debugger = None
        self._state_cache = None
        self._debugger = None
        self.tracker = Tracker()

    def __repr__(self):
        return f"<{self.name} device (wires={None if self.wires is None else len(self.wires)})>"

    # ---- capability queries ---------------------------------------------------------------
    def supports_derivatives(self, execution_config: ExecutionConfig | None = None, circuit=None):
        if execution_config is None:
            return True
        if execution_config.gradient_method not in ("adjoint", "best"):
            return False
        if circuit is None:
            return True
        try:
            self._adjoint_preprocess(circuit)
        except (DeviceError, QuantumFunctionError):
            return False
        return True

    supports_jvp = supports_derivatives
    supports_vjp = supports_derivatives

    def setup_execution_config(self, config: ExecutionConfig | None = None, circuit=None):
        config = config or ExecutionConfig()
        for option in config.device_options:
            if option not in self._device_options:
                raise DeviceError(f"device option {option} not present on {self}")
        updated = {}
        method = config.gradient_method
        if method == "best":
            method = "adjoint"
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
        opts.setdefault("exact_sampling", self._exact_sampling)
        opts.setdefault("fusion", self._fusion)
        opts.setdefault("return_torch", self._return_torch)
        updated["device_options"] = opts
        updated["mcm_config"] = self._setup_mcm_config(config.mcm_config, circuit)
        return replace(config, **updated)

    def _setup_mcm_config(self, mcm_config, tape):
        """default_qubit.py:739-760.  "deferred" is a transform above this boundary
        (``defer_measurements``): tapes that still hold a ``MidMeasure`` when they reach
        ``preprocess`` without shots and without tree-traversal are rejected there."""
        final = mcm_config.mcm_method
        if final is None:
            final = "one-shot" if getattr(tape, "shots", None) else "deferred"
        elif final == "device":
            final = "tree-traversal"
        supported = {"one-shot", "deferred", "tree-traversal"}
        if final not in supported:
            raise DeviceError(f"mcm_method {final} not supported on {self.name}. "
                              f"Supported methods are {supported}")
        if mcm_config.postselect_mode == "fill-shots" and final != "deferred":
            raise DeviceError(
                "Using postselect_mode='fill-shots' is only supported with mcm_method='deferred'.")
        return replace(mcm_config, mcm_method=final)

    # ---- preprocessing ----------------------------------------------------------------------
    def _validate(self, tape: QuantumScript):
        if self.wires is not None:
            extra = set(tape.wires) - set(self.wires)
            if extra:
                raise DeviceError(  # preprocess.py:111 validate_device_wires
                    f"Cannot run circuit(s) on {self.name} as they contain wires not found on "
                    f"the device: {extra}")
        for m in tape.measurements:
            if tape.shots and m.kind in ("state", "density_matrix", "purity", "vn_entropy",
                                         "mutual_info"):
                raise DeviceError(f"Measurement {m} not accepted with finite shots on {self.name}")
            if not tape.shots and m.kind in ("sample", "counts", "shadow", "shadow_expval"):
                raise DeviceError(f"Measurement {m} not accepted for analytic simulation on "
                                  f"{self.name}.")

    def _adjoint_preprocess(self, tape: QuantumScript) -> QuantumScript:
        """_add_adjoint_transforms, default_qubit.py:315-349."""
        if tape.shots:
            raise DeviceError("Finite shots are not supported with adjoint + b200.qubit")
        tape = _decompose(tape, adjoint_ops, "adjoint + b200.qubit")
        # default_qubit.py:243-283 (adjoint_state_measurements): all expectation values, or the
        # state itself (what the reference turns every other observable-free measurement into)
        state_only = len(tape.measurements) == 1 and tape.measurements[0].kind == "state"
        for m in tape.measurements:
            if state_only:
                break
            if m.kind != "expval":
                raise DeviceError(f"Measurement {m} not accepted with adjoint + b200.qubit "
                                  "(only expectation values, or the state alone).")
            if not adjoint_observables(m.obs):
                raise DeviceError(f"Observable {m.obs} not supported with adjoint + b200.qubit")
        n_op_params = sum(len(op.data) for op in tape.operations)
        if any(t >= n_op_params for t in tape.trainable_params):
            # preprocess.py:242 validate_adjoint_trainable_params
            raise QuantumFunctionError(
                "Differentiating with respect to the input parameters of an observable is not "
                "supported with the adjoint differentiation method.")
        return tape

    def preprocess(self, circuits, execution_config: ExecutionConfig | None = None):
        """Returns ``(tuple_of_tapes, config)`` — the composition of ``preprocess_transforms``
        and ``setup_execution_config`` (device_api.py:269-339)."""
        config = self.setup_execution_config(execution_config)
        single = isinstance(circuits, QuantumScript)
        tapes = [circuits] if single else list(circuits)
        out, posts = [], []
        for t in tapes:
            self._validate(t)
            t = _decompose(t, lambda op, _tr: stopping_condition(op), self.name)
            post = None
            tree = (execution_config is not None and execution_config.mcm_config is not None
                    and execution_config.mcm_config.mcm_method in ("tree-traversal", "device"))
            if tree and any(op.name == "MidMeasureMP" for op in t.operations):
                # default_qubit.py:660-661: the tape goes to simulate_tree_mcm as it is
                if config.gradient_method == "adjoint":
                    raise DeviceError("Mid-circuit measurements are not supported with adjoint + b200.qubit")
            elif any(op.name == "MidMeasureMP" for op in t.operations):
                # default_qubit.py:632-664 with mcm_method "one-shot" (the default with shots,
                # :744); the analytic default is "deferred", a transform above this boundary
                if not t.shots:
                    raise DeviceError(
                        "Mid-circuit measurements on b200.qubit run natively per shot "
                        "(mcm_method='one-shot') and need finite shots; apply defer_measurements "
                        "to the tape for analytic execution.")
                if config.gradient_method == "adjoint":
                    raise DeviceError("Finite shots are not supported with adjoint + b200.qubit")
                t, post = dynamic_one_shot(t)
            elif config.gradient_method == "adjoint":
                t = self._adjoint_preprocess(t)
            out.append(t)
            posts.append(post)
        return _PreprocessedBatch(out, posts), config

    # ---- execution -----------------------------------------------------------------------------
    def _as_batch(self, circuits):
        if isinstance(circuits, QuantumScript):
            return (circuits,), True
        return tuple(circuits), False

    def _track(self, circuits, kind, results=None):
        if not self.tracker.active:
            return
        if kind == "execute":
            self.tracker.update(batches=1)
            self.tracker.record()
            for c, r in zip(circuits, results):
                shots = c.shots.total_shots if c.shots else None
                bs = c.batch_size or 1
                if shots:
                    self.tracker.update(simulations=1, executions=bs, results=r, shots=shots * bs)
                else:
                    self.tracker.update(simulations=1, executions=bs, results=r)
                self.tracker.record()
        else:
            self.tracker.update(**{kind: 1})
            self.tracker.record()

    def _simulate(self, circuit, config):
        opts = (config.device_options if config else {}) or {}
        return _sim.simulate(
            circuit, rng=opts.get("rng", self._rng), dtype=opts.get("c_dtype", self._c_dtype),
            device=self._torch_device, exact_sampling=opts.get("exact_sampling", self._exact_sampling),
            state_cache=self._state_cache, fusion=opts.get("fusion", self._fusion),
            debugger=self._debugger, return_torch=bool(opts.get("return_torch", self._return_torch)),
            mcm_method=(config.mcm_config.mcm_method if config is not None and config.mcm_config else None))

    def execute(self, circuits, execution_config: ExecutionConfig | None = None):
        batch, single = self._as_batch(circuits)
        config = execution_config
        self._state_cache = {} if (config and config.use_device_jacobian_product) else None
        results = tuple(self._simulate(c, config) for c in batch)
        self._track(batch, "execute", results)
        return results[0] if single else results

    def _dtype(self, config):
        opts = (config.device_options if config else {}) or {}
        return opts.get("c_dtype", self._c_dtype)

    def _opt(self, config, name):
        opts = (config.device_options if config else {}) or {}
        return opts.get(name, getattr(self, f"_{name}"))

    def _fusion_of(self, config):
        """The ``fusion`` device option of the execution config, else the constructor's (the
        derivative entry points honour it like ``execute`` does)."""
        return int(self._opt(config, "fusion"))

    def compute_derivatives(self, circuits, execution_config: ExecutionConfig | None = None):
        batch, single = self._as_batch(circuits)
        self._track(batch, "derivative_batches")
        res = tuple(_adjoint.adjoint_jacobian(c, dtype=self._dtype(execution_config),
                                              device=self._torch_device, fusion=self._fusion_of(execution_config),
                                              return_torch=self._opt(execution_config, "return_torch"))
                    for c in batch)
        return res[0] if single else res

    def execute_and_compute_derivatives(self, circuits, execution_config=None):
        batch, single = self._as_batch(circuits)
        self._track(batch, "execute_and_derivative_batches")
        results, jacs = [], []
        for c in batch:
            c = c.map_to_standard_wires()
            jac, final = _adjoint.adjoint_jacobian(c, dtype=self._dtype(execution_config),
                                                   device=self._torch_device, return_state=True,
                                                   fusion=self._fusion_of(execution_config),
                                                   return_torch=self._opt(execution_config, "return_torch"))
            results.append(_sim.measure_final_state(c, final, False))
            jacs.append(jac)
        if single:
            return results[0], jacs[0]
        return tuple(results), tuple(jacs)

    def compute_jvp(self, circuits, tangents, execution_config=None):
        batch, single = self._as_batch(circuits)
        tangents = (tangents,) if single else tuple(tangents)
        self._track(batch, "jvp_batches")
        res = tuple(_adjoint.adjoint_jvp(c, t, dtype=self._dtype(execution_config),
                                         device=self._torch_device, fusion=self._fusion_of(execution_config))
                    for c, t in zip(batch, tangents))
        return res[0] if single else res

    def execute_and_compute_jvp(self, circuits, tangents, execution_config=None):
        res = self.execute(circuits, execution_config)
        return res, self.compute_jvp(circuits, tangents, execution_config)

    def compute_vjp(self, circuits, cotangents, execution_config=None):
        batch, single = self._as_batch(circuits)
        cotangents = (cotangents,) if single else tuple(cotangents)
        self._track(batch, "vjp_batches")
        def _state(circuit):            # default_qubit.py:1021-1029: reuse the forward state
            if not self._state_cache:
                return None
            return self._state_cache.get(circuit.map_to_standard_wires().hash)

        res = tuple(_adjoint.adjoint_vjp(c, t, dtype=self._dtype(execution_config),
                                         device=self._torch_device, fusion=self._fusion_of(execution_config),
                                         state=_state(c))
                    for c, t in zip(batch, cotangents))
        return res[0] if single else res

    def execute_and_compute_vjp(self, circuits, cotangents, execution_config=None):
        res = self.execute(circuits, execution_config)
        return res, self.compute_vjp(circuits, cotangents, execution_config)


def device(name: str = "b200.qubit", **kwargs) -> B200Qubit:
    """``qml.device(name, ...)`` for this package (devices/device_constructor.py:57)."""
    if name != "b200.qubit":
        raise DeviceError(f"Device {name} does not exist. This package provides 'b200.qubit'.")
    return B200Qubit(**kwargs)
