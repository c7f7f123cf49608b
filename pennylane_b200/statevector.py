"""Device-resident statevector: the host-side handle on the CUDA engine.

One ``StateVector`` owns a torch CUDA buffer of ``batch * 2**n`` complex amplitudes laid out
exactly like default.qubit's ``(B, 2, ..., 2)`` array flattened
(pennylane/devices/qubit/initialize_state.py:43-44), and forwards every numerical operation to
the C ABI in ``include/b200q.h``.  torch is used for memory, streams and (sharded mode)
``torch.distributed`` only; no torch arithmetic touches the amplitudes.

There is no CPU path: constructing a ``StateVector`` without a CUDA device, or without the built
``libb200q.so``, raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import B200QError, check, f64_array, int_array, u64_array
from .pauli import sentence_terms

_TORCH = None


def _torch():
    global _TORCH
    if _TORCH is None:
        import torch

        _TORCH = torch
    return _TORCH


_PHASE_GATES = {  # name -> phase applied to the |1> subspace of the (single) wire
    "PauliZ": -1.0 + 0j,
    "S": 1j,
    "T": np.exp(1j * np.pi / 4),
}
_CTRL_X_LIKE = {"CNOT": 1, "Toffoli": 2}


def _is_diagonal(mat: np.ndarray) -> bool:
    d = mat.shape[-1]
    off = mat[..., ~np.eye(d, dtype=bool)]
    return not np.any(off)


class StateVector:
    """``batch`` statevectors of ``num_wires`` qubits on one GPU.

    Args:
        num_wires: number of qubits (wire ``w`` is bit ``num_wires-1-w`` of the flat index).
        dtype: ``np.complex128`` (default, like default.qubit) or ``np.complex64``.
        batch: leading broadcast dimension (parameter broadcasting, simulate.py:211,235).
        device: torch device string.
    """

    def __init__(self, num_wires: int, dtype=np.complex128, batch: int = 1, device=None,
                 buffer=None):
        torch = _torch()
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise B200QError("pennylane_b200 needs a CUDA device (no CPU fallback exists)")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.n = int(num_wires)
        self.np_dtype = np.dtype(dtype)
        self.frozen = False                 # set on cached, never-again-written branch states
        self._cdf_cache = {}
        if self.np_dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise ValueError("dtype must be complex64 or complex128")
        self.dtype_code = 1 if self.np_dtype == np.complex128 else 0
        self.t_dtype = torch.complex128 if self.dtype_code else torch.complex64
        self.batch = int(batch)
        if buffer is not None:
            self.data = buffer
        else:
            self.data = torch.empty((self.batch, 1 << self.n), dtype=self.t_dtype, device=self.device)
            self.reset()
        self._work = None
        self._scal = torch.empty(4096, dtype=torch.float64, device=self.device)

    # ---- plumbing -----------------------------------------------------------------------
    @property
    def stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    @property
    def ptr(self):
        return C.c_void_p(self.data.data_ptr())

    def workspace(self, min_bytes: int = 0):
        torch = _torch()
        need = max(int(self.lib.b200q_workspace_bytes()), int(min_bytes))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.uint8, device=self.device)
        return C.c_void_p(self._work.data_ptr()), C.c_size_t(self._work.numel())

    def bit(self, wire: int) -> int:
        return self.n - 1 - int(wire)

    def bits(self, wires: Sequence[int]):
        return [self.n - 1 - int(w) for w in wires]

    def _upload(self, arr: np.ndarray):
        torch = _torch()
        arr = np.ascontiguousarray(arr, dtype=self.np_dtype)
        return torch.from_numpy(arr).to(self.device, non_blocking=False)

    # ---- state management ------------------------------------------------------------------
    def reset(self, index: int = 0):
        check(self.lib.b200q_set_basis_state(self.ptr, self.n, self.dtype_code, self.batch,
                                             int(index), self.stream))

    def set_state(self, state: np.ndarray):
        """Load amplitudes from the host: shape (2**n,), (2,)*n, or with a leading batch axis."""
        torch = _torch()
        state = np.asarray(state)
        dim = 1 << self.n
        flat = state.reshape(-1, dim) if state.size != dim else state.reshape(1, dim)
        if flat.shape[0] != self.batch:
            self._resize_batch(flat.shape[0], copy=False)
        self.data.copy_(torch.from_numpy(np.ascontiguousarray(flat, dtype=self.np_dtype)))

    def _resize_batch(self, batch: int, copy: bool = True):
        torch = _torch()
        if batch == self.batch:
            return
        if self.batch != 1 and copy:
            raise ValueError(f"cannot broadcast a batch of {self.batch} states to {batch}")
        new = torch.empty((batch, 1 << self.n), dtype=self.t_dtype, device=self.device)
        if copy:
            new.copy_(self.data.expand(batch, -1))
        self.data = new
        self.batch = batch

    def clone(self) -> "StateVector":
        return StateVector(self.n, self.np_dtype, self.batch, self.device, buffer=self.data.clone())

    def to_numpy(self) -> np.ndarray:
        """Host copy shaped like default.qubit's state: (2,)*n, or (B, 2, ..., 2) if batched."""
        out = self.data.cpu().numpy()
        shape = (2,) * self.n
        return out.reshape((self.batch,) + shape) if self.batch > 1 else out.reshape(shape)

    # ---- raw kernel wrappers ----------------------------------------------------------------
    def apply_matrix(self, mat: np.ndarray, wires: Sequence[int], control_wires: Sequence[int] = (),
                     control_values: Sequence[int] | None = None):
        """Dense (2^k x 2^k, optionally batched) matrix on ``wires`` controlled on
        ``control_wires``."""
        mat = np.asarray(mat)
        k = len(wires)
        d = 1 << k
        batched = mat.ndim == 3
        if batched and mat.shape[0] != self.batch:
            self._resize_batch(mat.shape[0])
        cvals = [1] * len(control_wires) if control_values is None else [int(bool(v)) for v in control_values]
        tgt = int_array(self.bits(wires))
        cb = int_array(self.bits(control_wires))
        cv = int_array(cvals)
        if not batched and k <= 3:
            host = np.ascontiguousarray(mat, dtype=np.complex128)
            check(self.lib.b200q_apply_matrix(self.ptr, self.n, self.dtype_code, self.batch, tgt, k,
                                              cb, cv, len(cvals), host.ctypes.data_as(C.c_void_p),
                                              None, 0, self.stream))
            return
        dev = self._upload(mat)
        check(self.lib.b200q_apply_matrix(self.ptr, self.n, self.dtype_code, self.batch, tgt, k, cb,
                                          cv, len(cvals), None, C.c_void_p(dev.data_ptr()),
                                          d * d if batched else 0, self.stream))
        self._keepalive = dev

    def apply_diag(self, diag: np.ndarray, wires: Sequence[int]):
        diag = np.asarray(diag)
        k = len(wires)
        batched = diag.ndim == 2
        if batched and diag.shape[0] != self.batch:
            self._resize_batch(diag.shape[0])
        bits = int_array(self.bits(wires))
        if not batched and k <= 6:
            host = np.ascontiguousarray(diag, dtype=np.complex128)
            check(self.lib.b200q_apply_diag(self.ptr, self.n, self.dtype_code, self.batch, bits, k,
                                            host.ctypes.data_as(C.c_void_p), None, 0, self.stream))
            return
        dev = self._upload(diag)
        check(self.lib.b200q_apply_diag(self.ptr, self.n, self.dtype_code, self.batch, bits, k, None,
                                        C.c_void_p(dev.data_ptr()), (1 << k) if batched else 0,
                                        self.stream))
        self._keepalive = dev

    def apply_phase(self, phase, control_wires: Sequence[int] = (), control_values=None):
        """Multiply the subspace where ``control_wires == control_values`` by ``phase``
        (scalar, or one value per batch element)."""
        phase = np.asarray(phase, dtype=np.complex128)
        cvals = [1] * len(control_wires) if control_values is None else [int(bool(v)) for v in control_values]
        cb = int_array(self.bits(control_wires))
        cv = int_array(cvals)
        if phase.ndim == 0:
            check(self.lib.b200q_apply_phase(self.ptr, self.n, self.dtype_code, self.batch, cb, cv,
                                             len(cvals), float(phase.real), float(phase.imag), None,
                                             self.stream))
            return
        if phase.shape[0] != self.batch:
            self._resize_batch(phase.shape[0])
        dev = self._upload(phase)
        check(self.lib.b200q_apply_phase(self.ptr, self.n, self.dtype_code, self.batch, cb, cv,
                                         len(cvals), 0.0, 0.0, C.c_void_p(dev.data_ptr()),
                                         self.stream))
        self._keepalive = dev

    def apply_parity_phase(self, theta, wires: Sequence[int]):
        """exp(-i theta/2 Z..Z) on ``wires``."""
        theta = np.asarray(theta, dtype=np.float64)
        mask = 0
        for b in self.bits(wires):
            mask |= 1 << b
        p0, p1 = np.exp(-0.5j * theta), np.exp(0.5j * theta)
        if theta.ndim == 0:
            check(self.lib.b200q_apply_parity_phase(self.ptr, self.n, self.dtype_code, self.batch,
                                                    mask, p0.real, p0.imag, p1.real, p1.imag, None,
                                                    self.stream))
            return
        if theta.shape[0] != self.batch:
            self._resize_batch(theta.shape[0])
        dev = self._upload(np.stack([p0, p1], axis=-1))
        check(self.lib.b200q_apply_parity_phase(self.ptr, self.n, self.dtype_code, self.batch, mask,
                                                0, 0, 0, 0, C.c_void_p(dev.data_ptr()), self.stream))
        self._keepalive = dev

    def apply_pauli_rot(self, theta, pauli_word: str, wires: Sequence[int]):
        """exp(-i theta/2 P) for any Pauli word, one sweep."""
        active = [(c, w) for c, w in zip(pauli_word, wires) if c != "I"]
        if not active:
            self.apply_phase(np.exp(-0.5j * np.asarray(theta)))
            return
        if all(c == "Z" for c, _ in active):
            self.apply_parity_phase(theta, [w for _, w in active])
            return
        xm = zm = ny = 0
        for c, w in active:
            b = 1 << self.bit(w)
            if c in "XY":
                xm |= b
            if c in "ZY":
                zm |= b
            ny += c == "Y"
        theta = np.asarray(theta, dtype=np.float64)
        cs, sn = np.cos(theta / 2), np.sin(theta / 2)
        if theta.ndim == 0:
            check(self.lib.b200q_apply_pauli_rot(self.ptr, self.n, self.dtype_code, self.batch, xm,
                                                 zm, ny, float(cs), float(sn), None, self.stream))
            return
        if theta.shape[0] != self.batch:
            self._resize_batch(theta.shape[0])
        dev = self._upload(cs + 1j * sn)
        check(self.lib.b200q_apply_pauli_rot(self.ptr, self.n, self.dtype_code, self.batch, xm, zm,
                                             ny, 0.0, 0.0, C.c_void_p(dev.data_ptr()), self.stream))
        self._keepalive = dev

    # ---- mid-circuit measurement (apply_operation.py:355-497) ---------------------------------
    def collapse(self, wire: int, sample: int, reset: bool, scale: float):
        """Project ``wire`` on ``sample``, multiply the survivors by ``scale`` and optionally
        reset the wire to |0> — one sweep (``b200q_collapse``)."""
        check(self.lib.b200q_collapse(self.ptr, self.n, self.dtype_code, self.bit(wire),
                                      int(sample), int(bool(reset)), float(scale), self.stream))

    def apply_mid_measure(self, op, mid_measurements: dict, rng=None):
        """``apply_mid_measure`` (apply_operation.py:415-497): sample the wire with the host
        Generator (``rng.binomial(1, 1 - p0)``, the same draw the reference makes), record it in
        ``mid_measurements`` and collapse + renormalise + reset on the device."""
        if self.batch > 1:
            raise ValueError("MidMeasure cannot be applied to batched states.")
        if mid_measurements is None:
            raise AssertionError("mid_measurements dictionary is required for MidMeasure")
        wire = op.wires[0]
        p = self.probs([wire])                                   # one read sweep: (p0, p1)
        sample, scale = self.mid_measure_draw(p, self.np_dtype, rng)
        mid_measurements[op] = sample
        self.collapse(wire, sample, bool(getattr(op, "reset", False)), scale)
        return sample

    @staticmethod
    def mid_measure_draw(p, np_dtype, rng=None):
        """The host half of ``apply_mid_measure`` (apply_operation.py:450-485) from the measured
        wire's marginal ``p = (p0, p1)``: the reference's checks, its one ``binomial(1, 1 - p0)``
        draw, and the factor ``1 / ||P psi||`` the collapse applies."""
        # :450 prob0 = real(norm(slice))**2 — sqrt then square, like the reference
        prob0 = float(np.sqrt(p[0])) ** 2
        eps = 10 * np.finfo(np_dtype).eps                        # :452-457
        if (prob0 - 1) > eps:
            raise ValueError(f"probabilities greater than 1. Got norm {prob0}.")
        if prob0 > 1:
            prob0 = prob0 / prob0
        binomial = np.random.binomial if rng is None else rng.binomial
        sample = int(binomial(1, 1 - prob0))
        # :478-485 projector then state / norm(state); numpy's complex / real multiplies by the
        # reciprocal.  A zero-probability branch (cannot be drawn by a Bernoulli with p = 0 or 1)
        # would give inf, as 0 / 0 does in the reference.
        norm = float(np.sqrt(p[sample]))
        with np.errstate(divide="ignore"):
            scale = float(np.float64(1.0) / np.float64(norm))
        return sample, scale

    # ---- operator dispatch (apply_operation.py:258-351 singledispatch, re-done for kernels) --
    def apply_operation(self, op, mid_measurements=None, rng=None):
        """Apply one operator (ours or a duck-typed PennyLane one) in place."""
        name = op.name
        wires = list(op.wires)
        if name == "MidMeasureMP":
            self.apply_mid_measure(op, mid_measurements, rng)
            return
        if name.startswith("Conditional") and hasattr(op, "meas_val"):
            # apply_conditional, apply_operation.py:355-411
            if op.meas_val.concretize(mid_measurements):
                self.apply_operation(op.base, mid_measurements=mid_measurements, rng=rng)
            return
        SWEEPS["per_gate"] += 1
        if name in ("Identity", "Barrier", "WireCut", "Snapshot"):
            return
        if name == "GlobalPhase":
            self.apply_phase(np.exp(-1j * np.asarray(op.data[0], dtype=np.float64)))
            return
        if name in _PHASE_GATES:
            self.apply_phase(_PHASE_GATES[name], wires)
            return
        if name in ("PhaseShift", "U1"):
            self.apply_phase(np.exp(1j * np.asarray(op.data[0], dtype=np.float64)), wires)
            return
        if name in ("CZ", "CCZ"):
            self.apply_phase(-1.0, wires)
            return
        if name == "ControlledPhaseShift":
            self.apply_phase(np.exp(1j * np.asarray(op.data[0], dtype=np.float64)), wires)
            return
        if name in ("RZ", "IsingZZ", "MultiRZ"):
            self.apply_parity_phase(op.data[0], wires)
            return
        if name == "PauliRot":
            self.apply_pauli_rot(op.data[0], op.hyperparameters["pauli_word"], wires)
            return
        if name in _CTRL_X_LIKE:
            nc = _CTRL_X_LIKE[name]
            self.apply_matrix(_XMAT, wires[nc:], wires[:nc])
            return
        if name == "MultiControlledX":
            cvals = op.hyperparameters.get("control_values")
            if cvals is None:
                cvals = getattr(op, "control_values", None)
            self.apply_matrix(_XMAT, wires[-1:], wires[:-1], cvals)
            return
        base = getattr(op, "base", None)
        cw = list(getattr(op, "control_wires", ()) or ())
        if base is not None and cw and (name.startswith("C(") or name == "ControlledQubitUnitary"):
            cvals = getattr(op, "control_values", None)
            tw = [w for w in wires if w not in cw]
            if getattr(base, "has_matrix", True):
                m = np.asarray(base.matrix())
                if len(tw) and m.shape[-1] == (1 << len(tw)):
                    bw = list(base.wires)
                    self._apply_matrix_auto(m, bw, cw, cvals)
                    return
        if name in ("CRX", "CRY", "CRZ", "CRot", "CY", "CH", "CSWAP") and cw:
            # textbook controlled gates: apply the base block to the target subspace only
            m = np.asarray(op.matrix())
            tw = [w for w in wires if w not in cw]
            d = 1 << len(tw)
            self._apply_matrix_auto(m[..., -d:, -d:], tw, cw, None)
            return
        if not getattr(op, "has_matrix", True):
            raise B200QError(f"operator {name} has no matrix and no dedicated kernel")
        self._apply_matrix_auto(np.asarray(op.matrix()), wires, (), None)

    # ---- fused execution (host fusion pass + tile kernels) ------------------------------------
    def rt_geometry(self, nvec: int = 1):
        """(T, RB, threads) of the register-tiled kernel for this dtype (b200q_rtile_geometry)."""
        if self.jit_enabled(nvec):
            from . import segjit

            g = segjit.default_geometry(self.dtype_code, nvec)
            return g.T, g.RB, 1 << g.TB
        key = (self.dtype_code, nvec)
        if key not in _RT_GEOM:
            T, RB, th = C.c_int(), C.c_int(), C.c_int()
            check(self.lib.b200q_rtile_geometry(self.dtype_code, nvec, C.byref(T), C.byref(RB),
                                                C.byref(th)))
            _RT_GEOM[key] = (T.value, RB.value, th.value)
        return _RT_GEOM[key]

    def default_tile(self, nvec: int = 1):
        """(T, L): tile bits / contiguous low bits.  States with at least T qubits use the
        register-tiled kernel (T fixed by its geometry); smaller states the shared-memory one."""
        import os

        T, _, _ = self.rt_geometry(nvec)
        if self.n < T:
            T = min(self.n, 12 if self.dtype_code else 13)
        # contiguous run of 2^L amplitudes per bulk copy / warp store: 256-byte runs reach the same
        # copy ceiling as 512-byte ones (measured 5.86 vs 5.95 TB/s) and free one more tile bit
        L = int(os.environ.get("B200Q_TILE_L", 4 if (self.dtype_code and T <= 11) else 5))
        return T, min(L, T)

    def jit_enabled(self, nvec: int = 1) -> bool:
        """Whether fused segments go through the structure-specialised kernels (segjit.py /
        csrc/segk.cuh) instead of the record interpreter.  ``B200Q_JIT``: 1 = always, 0 = never,
        unset = for states of at least ``B200Q_JIT_MIN_QUBITS`` (default 22) qubits, where the
        one-off NVRTC compilation of a structure (about a second) is small against the sweeps."""
        import os

        from . import segjit

        mode = os.environ.get("B200Q_JIT", "auto")
        geom = segjit.default_geometry(self.dtype_code, nvec)
        if mode == "0" or self.n < geom.T:
            return False
        if mode == "auto" and not getattr(self, "_jit_hot", False) \
                and self.n < int(os.environ.get("B200Q_JIT_MIN_QUBITS", 22)):
            return False
        return bool(self.lib.b200q_jit_available())

    def jit_possible(self, nvec: int = 1) -> bool:
        """The specialised kernels could run on this state (ignoring the size threshold): used by
        the program cache to promote a circuit STRUCTURE that is executed a second time — a
        training loop on a small register — to the compiled path."""
        import os

        from . import segjit

        if os.environ.get("B200Q_JIT", "auto") == "0":
            return False
        return self.n >= segjit.default_geometry(self.dtype_code, nvec).T \
            and bool(self.lib.b200q_jit_available())

    class _Hot:
        """``with sv.hot():`` — run the specialised path regardless of the size threshold."""

        def __init__(self, sv):
            self.sv = sv

        def __enter__(self):
            self.prev = getattr(self.sv, "_jit_hot", False)
            self.sv._jit_hot = True

        def __exit__(self, *a):
            self.sv._jit_hot = self.prev

    def hot(self):
        return StateVector._Hot(self)

    def compile_fused(self, ops_, level: int = 1, T: int | None = None, L: int | None = None,
                      bit_of=None):
        """Operators -> segments for this state (host fusion pass, compiler.py)."""
        import os

        from .compiler import compile_ops

        dT, dL = self.default_tile()
        rtT, _, _ = self.rt_geometry(1)
        jit = self.jit_enabled(1) and (T or dT) == rtT
        if jit:
            from . import segjit

            dL = min(segjit.default_low_bits(self.dtype_code, 1), dT)
        budget = None
        if jit:
            budget = int(os.environ.get("B200Q_ROUND_BUDGET", 0)) or None      # tuning knob
        return compile_ops(ops_, self.n, bit_of=bit_of, level=level, T=T or dT, L=L if L is not None else dL,
                           batched_ok=(T or dT) == rtT and self.n >= rtT, fold_cx=not jit,
                           round_budget=budget, RB=self.rt_geometry(1)[1], sww=3 if self.dtype_code else 4)

    def apply_operations_fused(self, ops_, level: int = 1, T: int | None = None,
                               L: int | None = None, bit_of=None):
        """Apply a list of operators through the fusion pass (compiler.py): returns the number of
        state sweeps (kernel launches over the full state) that were issued."""
        from . import program

        prog, _ = program.get_program(self, ops_, level, T, L, bit_of)
        if prog is not None:
            # cached structure, tables bound to the current parameter values
            for seg, plan, tab in zip(prog.segs, prog.plans, prog.tables):
                if plan is None:
                    self.run_segment(seg)
                else:
                    self._launch_plan(plan, tab)
            return len(prog.segs)
        segs = self.compile_fused(ops_, level, T, L, bit_of)
        self.prepare_segments(segs)
        for seg in segs:
            self.run_segment(seg)
        return len(segs)

    def _launch_plan(self, plan, coefs, base_hi: int = 0, fix_mask: int = 0, fix_val: int = 0):
        from . import segjit

        if not fix_mask or not fix_val:          # a partial launch counts once (its first piece)
            SWEEPS["fused_segments"] += 1

        if coefs.ndim == 2 and coefs.shape[0] != self.batch:   # broadcast parameters
            if self.batch != 1:
                raise ValueError(f"broadcast gates of batch {coefs.shape[0]} on a state of "
                                 f"batch {self.batch}")
            self._resize_batch(coefs.shape[0])
        w, wb = self.workspace()
        segjit.launch(plan, coefs, self.ptr, None, self.n, self.batch, w, wb, self.stream,
                      base_hi=base_hi, fix_mask=fix_mask, fix_val=fix_val)

    def prepare_segments(self, segs):
        """Plan every tile segment for the specialised kernels and compile the structures that
        are in no cache yet, in parallel (no-op on the interpreter path)."""
        from . import segjit

        if not self.jit_enabled(1):
            return
        rtT = self.rt_geometry(1)[0]
        geom = segjit.default_geometry(self.dtype_code, 1)
        plans = []
        for seg in segs:
            if seg.tile_bits is None or len(seg.tile_bits) != rtT:
                continue
            if getattr(seg, "_sk_plan", None) is None:
                seg._sk_plan = segjit.plan_segment(seg, geom, _low_run(seg.tile_bits))
                seg._sk_coefs = segjit.coefficients(seg._sk_plan, seg.prims)
            plans.append(seg._sk_plan)
        segjit.ensure_compiled(plans)

    def _run_segment_jit(self, seg, base_hi: int = 0, fix_mask: int = 0, fix_val: int = 0):
        """One fused segment through its structure-specialised kernel.  The plan (structure) and
        the coefficient table (values) are kept on the segment object.  ``fix_mask`` /
        ``fix_val``: partial launch over the tiles whose non-tile index bits ``fix_mask`` equal
        ``fix_val`` (the sharded engine pipelines segments against the exchange piece by piece)."""
        from . import segjit

        plan = getattr(seg, "_sk_plan", None)
        if plan is None:
            geom = segjit.default_geometry(self.dtype_code, 1)
            plan = segjit.plan_segment(seg, geom, _low_run(seg.tile_bits))
            seg._sk_plan = plan
            seg._sk_coefs = segjit.coefficients(plan, seg.prims)
        self._launch_plan(plan, seg._sk_coefs, base_hi, fix_mask, fix_val)

    def segment_partial_ok(self, seg) -> bool:
        """Whether :meth:`run_segment` can run ``seg`` on a subset of its tiles."""
        return seg.tile_bits is not None and self.jit_enabled(1) \
            and len(seg.tile_bits) == self.rt_geometry(1)[0]

    def run_segment(self, seg, base_hi: int = 0, fix_mask: int = 0, fix_val: int = 0):
        from .compiler import DIAG, encode_rt_segment, encode_segment

        if seg.tile_bits is not None and self.jit_enabled(1) \
                and len(seg.tile_bits) == self.rt_geometry(1)[0]:
            return self._run_segment_jit(seg, base_hi, fix_mask, fix_val)
        if fix_mask:
            raise B200QError("partial launches exist for the specialised segment kernels only")
        SWEEPS["fused_segments"] += seg.tile_bits is not None
        if seg.tile_bits is None:
            p = seg.prims[0]
            if p.op is not None:
                self.apply_operation(p.op)
            elif p.kind == DIAG:
                self.apply_diag(np.asarray(p.mat), [self.n - 1 - b for b in p.other])
            else:  # pragma: no cover
                raise B200QError("unexpected generic primitive")
            return
        w, wb = self.workspace()
        rtT, rtRB, _ = self.rt_geometry(1)
        if len(seg.tile_bits) == rtT:
            enc = getattr(seg, "_rt_enc", None)
            if enc is None:
                enc = encode_rt_segment(seg, rtRB, 3 if self.dtype_code else 4)
                seg._rt_enc = enc
            ops_arr, table, nrec = enc
            if table.ndim == 2:                    # broadcast parameters: one table per batch element
                if table.shape[0] != self.batch:
                    if self.batch != 1:
                        raise ValueError(f"broadcast gates of batch {table.shape[0]} on a state of "
                                         f"batch {self.batch}")
                    self._resize_batch(table.shape[0])
                fn = self.lib.b200q_apply_rtile_bcast
            else:
                fn = self.lib.b200q_apply_rtile
            check(fn(
                self.ptr, None, self.n, self.dtype_code, self.batch, int_array(seg.tile_bits),
                rtT, _low_run(seg.tile_bits), C.cast(ops_arr, C.c_void_p), nrec,
                table.ctypes.data_as(C.c_void_p), int(table.shape[-1]), 0, 1, int(base_hi), 1.0, None,
                w, wb, self.stream))
            return
        ops_arr, table = encode_segment(seg)
        check(self.lib.b200q_apply_tile(
            self.ptr, self.n, self.dtype_code, self.batch, int_array(seg.tile_bits),
            len(seg.tile_bits), _low_run(seg.tile_bits),
            C.cast(ops_arr, C.c_void_p), len(ops_arr), table.ctypes.data_as(C.c_void_p),
            int(table.size), w, wb, self.stream))

    def _apply_matrix_auto(self, mat, wires, cw, cvals):
        if _is_diagonal(mat) and not cw:
            idx = np.arange(mat.shape[-1])
            self.apply_diag(mat[..., idx, idx], wires)
        else:
            self.apply_matrix(mat, wires, cw, cvals)

    # ---- measurements -----------------------------------------------------------------------
    def probs(self, wires: Sequence[int] | None = None) -> np.ndarray:
        """Marginal probabilities over ``wires`` in the given order (probs.py:101-135)."""
        return self.probs_device(wires).cpu().numpy().reshape(
            (self.batch, -1) if self.batch > 1 else (-1,))

    def probs_device(self, wires: Sequence[int] | None = None):
        torch = _torch()
        wires = list(range(self.n)) if wires is None else list(wires)
        m = len(wires)
        out = torch.empty((self.batch, 1 << m), dtype=torch.float64, device=self.device)
        w, wb = self.workspace()
        check(self.lib.b200q_probs(self.ptr, self.n, self.dtype_code, self.batch,
                                   int_array(self.bits(wires)), m, C.c_void_p(out.data_ptr()), w, wb,
                                   self.stream))
        return out

    def reduced_dm(self, wires: Sequence[int], fixed_zero: Sequence[int] = ()) -> np.ndarray:
        """Reduced density matrix over ``wires`` (in that order), complex128, shape
        ``(2^m, 2^m)`` or ``(batch, 2^m, 2^m)`` — ``reduce_statevector``
        (pennylane/math/quantum.py:386-487) without ever forming more than the result.

        The last two wires are the in-thread block of ``b200q_gram_block``; the others are
        enumerated here as (row, column >= row) outer assignments, the lower triangle follows
        from Hermiticity.  ``fixed_zero``: wires known to be |0> (already measured and reset);
        only the 2^-len(fixed_zero) of the state where they are 0 is read."""
        torch = _torch()
        wires = list(wires)
        m = len(wires)
        fixed = list(fixed_zero)
        if len(set(wires + fixed)) != m + len(fixed) or any(w < 0 or w >= self.n for w in wires + fixed):
            raise ValueError(f"reduced_dm: bad wires {wires} / {fixed}")
        mi = min(m, 2)
        mo = m - mi
        outer, inner = wires[:mo], wires[mo:]
        # fixed wires are the most significant outer bits: assignments below 2^mo leave them 0
        ib, ob = int_array(self.bits(inner)), int_array(self.bits(fixed + outer))
        n_outer = mo + len(fixed)
        D = 1 << mi
        pairs = [(a, b) for a in range(1 << mo) for b in range(a, 1 << mo)]
        out = torch.empty((self.batch, len(pairs), 2 * D * D), dtype=torch.float64,
                          device=self.device)
        w, wb = self.workspace()
        esz = self.np_dtype.itemsize
        for bi in range(self.batch):
            base = self.data.data_ptr() + bi * (esz << self.n)
            for k, (a, b) in enumerate(pairs):
                check(self.lib.b200q_gram_block(
                    C.c_void_p(base), self.n, self.dtype_code, ib, mi, ob, n_outer, a, b,
                    C.c_void_p(out[bi, k].data_ptr()), w, wb, self.stream))
        blocks = out.cpu().numpy().reshape(self.batch, len(pairs), D, D, 2)
        blocks = blocks[..., 0] + 1j * blocks[..., 1]
        rho = np.empty((self.batch, 1 << mo, D, 1 << mo, D), dtype=np.complex128)
        for k, (a, b) in enumerate(pairs):
            if a == b:
                # the kernel accumulates only the upper triangle of a diagonal block (with a real
                # diagonal): mirror it, so the result is exactly Hermitian
                up = np.triu(blocks[:, k], 1)
                blocks[:, k] = np.triu(blocks[:, k]) + np.conj(np.swapaxes(up, -1, -2))
            rho[:, a, :, b, :] = blocks[:, k]
            if a != b:
                rho[:, b, :, a, :] = np.conj(np.swapaxes(blocks[:, k], -1, -2))
        rho = rho.reshape(self.batch, 1 << m, 1 << m)
        return rho if self.batch > 1 else rho[0]

    def expval_csr(self, H, wires: Sequence[int]) -> np.ndarray:
        """<psi| H |psi> per batch element for a scipy CSR matrix ``H`` (2^k x 2^k) on ``wires``
        (measure.py:74-118, SparseHamiltonian) — H is NOT expanded to the full register."""
        import scipy.sparse as sp

        torch = _torch()
        H = sp.csr_matrix(H)
        k = len(wires)
        if H.shape != (1 << k, 1 << k):
            raise ValueError(f"sparse matrix of shape {H.shape} does not act on {k} wires")
        H.sort_indices()
        dev = self.device
        indptr = torch.from_numpy(np.ascontiguousarray(H.indptr, dtype=np.int64)).to(dev)
        indices = torch.from_numpy(np.ascontiguousarray(H.indices, dtype=np.int64)).to(dev)
        data = torch.from_numpy(np.ascontiguousarray(H.data, dtype=np.complex128)).to(dev)
        w, wb = self.workspace()
        check(self.lib.b200q_expval_csr(
            self.ptr, self.n, self.dtype_code, self.batch, int_array(self.bits(wires)), k,
            C.c_void_p(indptr.data_ptr()), C.c_void_p(indices.data_ptr()), C.c_void_p(data.data_ptr()),
            C.c_void_p(self._scal.data_ptr()), w, wb, self.stream))
        return self._scal[: self.batch].cpu().numpy().copy()

    def expval_pauli_sentence(self, ps, wire_map=None) -> np.ndarray:
        """<psi| sum_t c_t P_t |psi> for a Pauli sentence (mapping word -> coeff).
        Returns a float (or (B,) array when batched)."""
        wire_to_bit = {w: self.bit(w if wire_map is None else wire_map[w])
                       for pw in ps for w in pw}
        xs, zs, ys, cs = sentence_terms(ps, wire_to_bit)
        cs = [float(np.real(c)) for c in cs]
        w, wb = self.workspace()
        check(self.lib.b200q_expval_pauli_sum(
            self.ptr, self.n, self.dtype_code, self.batch, u64_array(xs), u64_array(zs),
            int_array(ys) if ys else None, f64_array(cs), len(cs),
            C.c_void_p(self._scal.data_ptr()), w, wb, self.stream))
        res = self._scal[: self.batch].cpu().numpy().copy()
        return res if self.batch > 1 else res[0]

    def inner(self, other: "StateVector") -> np.ndarray:
        """<self|other> per batch element."""
        w, wb = self.workspace()
        check(self.lib.b200q_inner(self.ptr, other.ptr, self.n, self.dtype_code, self.batch,
                                   C.c_void_p(self._scal.data_ptr()), w, wb, self.stream))
        r = self._scal[: 2 * self.batch].cpu().numpy()
        res = r[: self.batch] + 1j * r[self.batch:]
        return res if self.batch > 1 else res[0]

    def norm2(self):
        return np.real(self.inner(self))

    def scale(self, factor):
        self.apply_phase(factor)

    def sample(self, shots: int, rng: np.random.Generator, wires: Sequence[int] | None = None,
               exact: bool = True) -> np.ndarray:
        """``shots`` samples of ``wires`` as a (shots, len(wires)) int64 array — or
        (B, shots, len(wires)) when batched — drawn exactly like sampling.py:500-531 with the
        HOST generator ``rng`` (one ``rng.random(shots)`` per batch element)."""
        torch = _torch()
        wires = list(range(self.n)) if wires is None else list(wires)
        m = len(wires)
        if shots == 0:                      # every shot postselected away (simulate.py:159-167)
            empty = np.zeros((0, m), dtype=np.int64)
            return np.stack([empty] * self.batch) if self.batch > 1 else empty
        # A frozen state (a cached branch of the one-shot MCM tree: never written again) keeps
        # the normalised CDF b200q_sample leaves behind; later draws are search + unpack only.
        frozen = getattr(self, "frozen", False) and self.batch == 1
        key = (tuple(wires), bool(exact))
        if frozen and key in self._cdf_cache:
            cdf, norm, has_nan = self._cdf_cache[key]
            u = torch.from_numpy(rng.random(shots)).to(self.device)
            if has_nan:
                return np.zeros((shots, m), dtype=np.int64)
            if abs(norm - 1.0) > 1e-6:
                raise ValueError("probabilities do not sum to 1")
            idx = torch.empty(shots, dtype=torch.int64, device=self.device)
            bits = torch.empty((shots, m), dtype=torch.int64, device=self.device)
            check(self.lib.b200q_search(C.c_void_p(cdf.data_ptr()), m, C.c_void_p(u.data_ptr()),
                                        shots, C.c_void_p(idx.data_ptr()), self.stream))
            check(self.lib.b200q_unpack_bits(C.c_void_p(idx.data_ptr()), shots, m,
                                             C.c_void_p(bits.data_ptr()), self.stream))
            return bits.cpu().numpy()
        probs = self.probs_device(wires)
        outs = []
        need = ((1 << m) // 128 + (1 << m) // (128 * 2047) + 128) * 8 + (4 << 20)
        w, wb = self.workspace(need)
        for b in range(self.batch):
            u = torch.from_numpy(rng.random(shots)).to(self.device)
            bits = torch.empty((shots, m), dtype=torch.int64, device=self.device)
            flags = torch.zeros(1, dtype=torch.int32, device=self.device)
            pb = probs[b]
            check(self.lib.b200q_sample(C.c_void_p(pb.data_ptr()), m, C.c_void_p(u.data_ptr()),
                                        shots, 0 if exact else 1, None,
                                        C.c_void_p(bits.data_ptr()),
                                        C.c_void_p(self._scal.data_ptr()),
                                        C.c_void_p(flags.data_ptr()), w, wb, self.stream))
            norm = float(self._scal[0].item())
            if frozen:
                self._cdf_cache[key] = (pb, norm, bool(int(flags.item())))
            if int(flags.item()):
                # sampling.py:322-325 — NaN probabilities give all-zero samples
                outs.append(np.zeros((shots, m), dtype=np.int64))
                continue
            if abs(norm - 1.0) > 1e-6:  # sampling.py:33, 514-519
                raise ValueError("probabilities do not sum to 1")
            outs.append(bits.cpu().numpy())
        return np.stack(outs) if self.batch > 1 else outs[0]


_XMAT = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_RT_GEOM: dict = {}
#: launches over the whole state issued by this process (tools/run_configs.py, bench.py): fused
#: segment launches and per-gate kernels
SWEEPS = {"fused_segments": 0, "per_gate": 0}



def _low_run(tile_bits) -> int:
    """Number of leading tile bits that are contiguous from bit 0 (the kernel's L)."""
    run = 0
    for i, b in enumerate(tile_bits):
        if b != i:
            break
        run += 1
    return run
