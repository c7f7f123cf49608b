"""CPU tests of the specialised-kernel host side (pennylane_b200/segjit.py): normalised forms,
plans emulated against the oracle, structure keys stable under parameter rebinding, and the
NVRTC compilation of generated kernels (NVRTC needs no GPU)."""
import numpy as np
import pytest

import pennylane_b200 as qb
from pennylane_b200 import compiler as cc
from pennylane_b200 import ops as q
from pennylane_b200 import segjit as sj

from sk_emulator import run_plan


def _random_unitary(rng):
    a = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    qm, r = np.linalg.qr(a)
    return qm * (np.diag(r) / np.abs(np.diag(r)))


def test_canon_reconstructs_standard_and_random_blocks():
    rng = np.random.default_rng(0)
    mats = [_random_unitary(rng) for _ in range(200)]
    for th in (0.0, 0.3, np.pi / 2, 2.9, np.pi, 4.0, 2 * np.pi):
        ry, rx, rz = (np.asarray(g(th, wires=0).matrix()) for g in (q.RY, q.RX, q.RZ))
        mats += [ry, rx, rz, rz @ ry, ry @ rz, rx @ rz, rz @ rx @ rz]
    mats += [np.asarray(q.Hadamard(wires=0).matrix()), np.array([[0, 1], [1, 0]], dtype=complex),
             np.array([[0, -1j], [1j, 0]]), np.eye(2, dtype=complex), np.diag([1, 1j])]
    for u in mats:
        c = sj.canon_1q(u)
        assert c is not None
        assert np.max(np.abs(c.matrix() - u)) < 5e-16 * max(1.0, np.max(np.abs(u))) * 4
        assert abs(c.t) <= 1.0 + 1e-15 and abs(abs(c.r) - 1) < 1e-14 and abs(abs(c.l) - 1) < 1e-14
        assert abs(c.s) >= 0.7                      # the pivot keeps the scalar away from zero


def test_canon_cheap_forms():
    """RY / RX: the kernel alone (4 FP64 per pair); RZ.RY and RY.RZ: one phase (8)."""
    for th in (0.3, 2.9, 4.0):
        c = sj.canon_1q(np.asarray(q.RY(th, wires=0).matrix()))
        assert (c.kern, c.dl, c.dr) == (0, False, False)
        c = sj.canon_1q(np.asarray(q.RX(th, wires=0).matrix()))
        assert (c.kern, c.dl, c.dr) == (1, False, False)
        rz, ry = np.asarray(q.RZ(0.7, wires=0).matrix()), np.asarray(q.RY(th, wires=0).matrix())
        c = sj.canon_1q(rz @ ry)
        assert (c.kern, c.dl, c.dr) == (0, True, False)
        c = sj.canon_1q(ry @ rz)
        assert (c.kern, c.dl, c.dr) == (0, False, True)


def test_non_unitary_block_is_not_normalised():
    assert sj.canon_1q(np.array([[1, 0], [0, 0]], dtype=complex)) is None
    assert sj.canon_1q(np.array([[1, 2], [3, 4]], dtype=complex)) is None


def _circuit(n, depth, seed):
    rng = np.random.default_rng(seed)
    ops_ = []
    for _ in range(depth):
        w = [int(x) for x in rng.permutation(n)]
        a, b, c = w[0], w[1], w[2]
        th = rng.uniform(0, 6)
        u2 = _random_unitary(rng)
        u4 = np.kron(_random_unitary(rng), _random_unitary(rng))
        u4 = u4 @ np.asarray(q.IsingXY(rng.uniform(0, 6), wires=[0, 1]).matrix())
        choices = [q.RX(th, wires=a), q.RY(th, wires=a), q.RZ(th, wires=a), q.PhaseShift(th, wires=a),
                   q.Hadamard(wires=a), q.PauliX(wires=a), q.PauliY(wires=a), q.PauliZ(wires=a),
                   q.S(wires=a), q.T(wires=a), q.QubitUnitary(u2, wires=a), q.Rot(th, 0.3 * th, 1.1, wires=a),
                   q.CNOT(wires=[a, b]), q.CZ(wires=[a, b]), q.CRY(th, wires=[a, b]), q.CRot(th, 1.0, 2.0, wires=[a, b]),
                   q.Toffoli(wires=[a, b, c]), q.SWAP(wires=[a, b]), q.CSWAP(wires=[a, b, c]),
                   q.IsingXX(th, wires=[a, b]), q.IsingZZ(th, wires=[a, b]), q.IsingXY(th, wires=[a, b]),
                   q.MultiRZ(th, wires=[a, b, c]), q.QubitUnitary(u4, wires=[a, b]),
                   q.ControlledPhaseShift(th, wires=[a, b]), q.GlobalPhase(th),
                   q.MultiControlledX(wires=[a, b, c], control_values=[0, 1]),
                   q.DiagonalQubitUnitary(np.exp(1j * rng.uniform(0, 6, size=8)), wires=[a, b, c])]
        ops_.append(choices[int(rng.integers(len(choices)))])
    return ops_


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("n,RB,TB,L", [(8, 3, 4, 3), (9, 3, 5, 4), (10, 4, 5, 5), (11, 4, 6, 4)])
def test_plans_reproduce_circuit(level, n, RB, TB, L):
    from oracle import simulate as o_sim

    ops_ = _circuit(n, 80, seed=10 * n + level)
    tape = qb.QuantumScript(ops_, [qb.state()])
    ref, _ = o_sim.get_final_state(tape)
    geom = sj.Geometry(1, RB, TB, 1, 2)
    segs = cc.compile_ops(ops_, n, level=level, T=geom.T, L=L, fold_cx=False)
    state = np.zeros(1 << n, dtype=complex)
    state[0] = 1.0
    ndk = 0
    for seg in segs:
        if seg.tile_bits is None:
            from oracle.apply_operation import apply_operation as o_apply
            p = seg.prims[0]
            if p.op is not None:
                state = o_apply(p.op, state.reshape((2,) * n)).reshape(-1)
            else:
                from test_compiler import _apply_prim
                state = _apply_prim(p, state, n)
            continue
        plan = sj.plan_segment(seg, geom, L)
        coefs = sj.coefficients(plan, seg.prims)
        ndk += sum(r[0] == "dk" for r in plan.ir)
        state = run_plan(plan, coefs, state, n)
    assert ndk > 0
    assert np.max(np.abs(state - ref.reshape(-1))) < 1e-12


def test_structure_key_is_independent_of_values_and_positions():
    """The 30-qubit ansatz compiles a handful of kernels; new angles reuse them."""
    import bench

    n = 30
    geom = sj.default_geometry(1, 1)
    keys = {}
    for seed in (3, 4):
        segs = cc.compile_ops(bench.hea_ops(n, seed=seed), n, level=1, T=geom.T, L=5, fold_cx=False)
        keys[seed] = [sj.plan_segment(s, geom, 5).key for s in segs]
    assert keys[3] == keys[4]
    assert len(set(keys[3])) <= 20
    # every block of the ansatz is RZ.RY: the real kernel + one phase
    segs = cc.compile_ops(bench.hea_ops(n), n, level=1, T=geom.T, L=5, fold_cx=False)
    plan = sj.plan_segment(segs[1], geom, 5)
    assert all(r[2:5] == (0, 1, 0) for r in plan.ir if r[0] == "dk")


def test_rebinding_with_special_values_raises_form_mismatch_only_when_needed():
    n, geom, L = 8, sj.Geometry(1, 3, 4, 1, 2), 3
    ops_a = [q.RY(0.3, wires=w) for w in range(n)]
    seg = cc.compile_ops(ops_a, n, level=1, T=geom.T, L=L, fold_cx=False)[0]
    plan = sj.plan_segment(seg, geom, L)
    # RY with other angles (any pivot) fits
    seg_b = cc.compile_ops([q.RY(3.0, wires=w) for w in range(n)], n, level=1, T=geom.T, L=L, fold_cx=False)[0]
    sj.coefficients(plan, seg_b.prims)
    # a general unitary in the same place does not
    seg_c = cc.compile_ops([q.Rot(0.3, 0.4, 0.5, wires=w) for w in range(n)], n, level=1, T=geom.T, L=L,
                           fold_cx=False)[0]
    with pytest.raises(sj.FormMismatch):
        sj.coefficients(plan, seg_c.prims)
    # planning with the old forms as a hint keeps the kernel when the values allow
    plan_b = sj.plan_segment(seg_b, geom, L, forms_hint=plan.forms)
    assert plan_b.key == plan.key


def test_generated_kernels_compile_with_nvrtc():
    """NVRTC needs no GPU: compile the kernels of a mixed circuit (forward) here."""
    import __graft_entry__ as g

    g.build()
    from pennylane_b200 import _lib

    lib = _lib.load()
    if not lib.b200q_jit_available():
        pytest.skip("libnvrtc not found")
    n = 14
    for dtype_code in (1, 0):
        geom = sj.default_geometry(dtype_code, 1)
        segs = cc.compile_ops(_circuit(n, 60, seed=5), n, level=1, T=geom.T, L=5, fold_cx=False)
        done = 0
        for seg in segs:
            if seg.tile_bits is None:
                continue
            plan = sj.plan_segment(seg, geom, 5)
            cubin = sj.compile_plan(plan)
            assert cubin[:4] == b"\x7fELF"
            done += 1
            if done == 2:
                break
        assert done


# ---- adjoint reverse sweep through the specialised plans (CPU, emulated) ------------------------
@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("n,RB,TB,L", [(8, 3, 4, 3), (9, 3, 5, 4), (10, 3, 6, 5)])
def test_reverse_sweep_plans_give_oracle_jacobian(level, n, RB, TB, L):
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from pennylane_b200.adjoint import _fused_reverse_program
    from test_compiler import _trainable_circuit

    ops_ = _trainable_circuit(n, 60, seed=n + level)
    obs = q.PauliZ(wires=0) @ q.PauliX(wires=2)
    tape = qb.QuantumScript(ops_, [qb.expval(obs)])
    state, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, state), dtype=float)
    prims, filled, trainable = _fused_reverse_program(tape, n, RB, level)
    geom = sj.Geometry(1, RB, TB, 2, 1)
    segs = cc.pack_segments(cc.merge_blocks(prims, level, fold_cx=False), n, T=geom.T, L=L, max_ops=12)
    _check_reverse(segs, geom, L, n, state, obs, ref, filled, trainable, expect_split=False)


def test_parameter_whose_generator_terms_straddle_segments():
    """the XX and YY terms of an IsingXY generator (or the I and Z terms of a
    PhaseShift) may land in different segments; the per-segment slot sums must be ADDED."""
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from pennylane_b200.adjoint import _fused_reverse_program

    n, RB, TB, L = 9, 3, 5, 4
    rng = np.random.default_rng(5)
    ops_ = []
    for _ in range(12):
        a, b = (int(x) for x in rng.permutation(n)[:2])
        ops_ += [q.RY(rng.uniform(0, 6), wires=a), q.IsingXY(rng.uniform(0, 6), wires=[a, b]),
                 q.PhaseShift(rng.uniform(0, 6), wires=b)]
    obs = q.PauliZ(wires=0) @ q.PauliY(wires=3)
    tape = qb.QuantumScript(ops_, [qb.expval(obs)])
    state, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, state), dtype=float)
    prims, filled, trainable = _fused_reverse_program(tape, n, RB, 1)
    geom = sj.Geometry(1, RB, TB, 2, 1)
    segs = cc.pack_segments(cc.merge_blocks(prims, 1, fold_cx=False), n, T=geom.T, L=L, max_ops=3)
    _check_reverse(segs, geom, L, n, state, obs, ref, filled, trainable, expect_split=True)


def _check_reverse(segs, geom, L, n, state, obs, ref, filled, trainable, expect_split):
    from oracle.apply_operation import apply_operation as o_apply
    from pennylane_b200.adjoint import _accumulate_slot_sums

    ket = state.reshape(-1).copy()
    bra = 2.0 * o_apply(obs, state).reshape(-1)
    raw, gather = [], []
    for seg in segs:
        assert seg.tile_bits is not None
        plan = sj.plan_segment(seg, geom, L)
        coefs = sj.coefficients(plan, seg.prims)
        ket, bra, sums = run_plan(plan, coefs, ket, n, bra=bra)
        for slot, param in enumerate(plan.slot_params):
            gather.append((len(raw) + slot, 0, param))
        raw += list(-sums)
    per_param = {}
    for _, _, param in gather:
        per_param[param] = per_param.get(param, 0) + 1
    if expect_split:
        assert max(per_param.values()) > 1
    jac = _accumulate_slot_sums(np.array(raw), gather, len(trainable), 1)[:, 0]
    assert sorted(filled) == list(range(len(trainable)))
    assert np.max(np.abs(jac - ref)) < 1e-12
    # both vectors carry the same gates: the ket is the initial state again
    init = np.zeros(1 << n, dtype=complex); init[0] = 1.0
    assert np.max(np.abs(ket - init)) < 1e-12


def test_reverse_sweep_of_the_ansatz_uses_one_block_per_wire_and_layer():
    import bench
    from pennylane_b200.adjoint import _fused_reverse_program

    n = 16
    tape = bench.hea_tape(n, 3)
    geom = sj.default_geometry(1, 2)
    prims, filled, trainable = _fused_reverse_program(tape, n, geom.RB, 1)
    merged = cc.merge_blocks(prims, 1, fold_cx=False)
    assert sum(p.kind == cc.DENSE1 for p in merged) == 3 * n
    segs = cc.pack_segments(merged, n, T=geom.T, L=5, max_ops=64)
    for seg in segs:
        plan = sj.plan_segment(seg, geom, 5)
        # RY^dagger RZ^dagger: the real kernel with one (right) phase
        assert all(r[2:5] == (0, 0, 1) for r in plan.ir if r[0] == "dk")


# ---- structure-keyed program cache with parameter rebinding (CPU) ------------------------------
class _FakeState:
    """The slice of StateVector the program cache reads (no CUDA)."""

    def __init__(self, n, geom, L, above_threshold=True):
        self.n, self.dtype_code, self._geom, self._L = n, geom.dtype_code, geom, L
        self.above_threshold, self._hot = above_threshold, False

    def jit_enabled(self, nvec=1):
        return self.above_threshold or self._hot

    def jit_possible(self, nvec=1):
        return True

    def hot(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self._hot = self._hot, True
            try:
                yield
            finally:
                self._hot = prev
        return ctx()

    def rt_geometry(self, nvec=1):
        return self._geom.T, self._geom.RB, 1 << self._geom.TB

    def default_tile(self, nvec=1):
        return self._geom.T, self._L

    def compile_fused(self, ops_, level=1, T=None, L=None, bit_of=None):
        return cc.compile_ops(ops_, self.n, bit_of=bit_of, level=level, T=T or self._geom.T,
                              L=self._L if L is None else L, batched_ok=True, fold_cx=False)

    def prepare_segments(self, segs):
        for seg in segs:
            if seg.tile_bits is not None:
                seg._sk_plan = sj.plan_segment(seg, self._geom, self._L)
                seg._sk_coefs = sj.coefficients(seg._sk_plan, seg.prims)


def _run_program(prog, n):
    from oracle.apply_operation import apply_operation as o_apply

    state = np.zeros(1 << n, dtype=complex)
    state[0] = 1.0
    for seg, plan, tab in zip(prog.segs, prog.plans, prog.tables):
        if plan is None:
            state = o_apply(seg.prims[0].op, state.reshape((2,) * n)).reshape(-1)
        else:
            state = run_plan(plan, tab, state, n)
    return state


def _param_circuit(n, rng, special=False):
    """Fixed structure, parameters drawn from ``rng`` (``special``: angles 0 / pi / pi/2)."""
    def ang():
        return float(rng.choice([0.0, np.pi, np.pi / 2, 2 * np.pi])) if special else float(rng.uniform(0, 6))
    ops_ = []
    for l in range(3):
        for w in range(n):
            ops_ += [q.RY(ang(), wires=w), q.RZ(ang(), wires=w)]
        ops_ += [q.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
        ops_ += [q.RX(ang(), wires=0), q.Hadamard(wires=1), q.PhaseShift(ang(), wires=2),
                 q.Rot(ang(), ang(), ang(), wires=3), q.CRY(ang(), wires=[4, 5]),
                 q.IsingZZ(ang(), wires=[5, 6]), q.IsingXY(ang(), wires=[6, 7]), q.RZ(ang(), wires=7),
                 q.T(wires=0), q.MultiRZ(ang(), wires=[0, 3, 5]), q.GlobalPhase(ang()),
                 q.QubitUnitary(_random_unitary(rng), wires=1)]
    return ops_


def test_program_cache_rebinds_parameters_without_recompiling(monkeypatch):
    from oracle import simulate as o_sim
    from pennylane_b200 import program

    program.clear()
    n, geom, L = 9, sj.Geometry(1, 3, 5, 1, 2), 4
    sv = _FakeState(n, geom, L)
    calls = {"n": 0}
    real = cc.compile_ops

    def counting(*a, **k):
        calls["n"] += 1
        return real(*a, **k)

    monkeypatch.setattr(cc, "compile_ops", counting)
    rng = np.random.default_rng(0)
    keys = None
    for it in range(4):
        ops_ = _param_circuit(n, rng, special=(it == 3))
        prog, hit = program.get_program(sv, ops_, 1)
        assert prog is not None
        ref, _ = o_sim.get_final_state(qb.QuantumScript(ops_, [qb.state()]))
        got = _run_program(prog, n)
        assert np.max(np.abs(got - ref.reshape(-1))) < 1e-12
        if it == 0:
            assert not hit and calls["n"] == 1
            keys = [p.key for p in prog.plans if p is not None]
        elif it < 3:
            # new generic angles: zero compile_ops calls, same kernels
            assert hit and calls["n"] == 1
            assert [p.key for p in prog.plans if p is not None] == keys
    assert program.STATS["hits"] >= 2


def test_program_cache_first_call_with_special_values_then_generic():
    """Compiled at angles 0 / pi (blocks happen to be diagonal / permutations), then generic
    angles: the entry is dropped and rebuilt instead of producing a wrong state."""
    from oracle import simulate as o_sim
    from pennylane_b200 import program

    program.clear()
    n, geom, L = 9, sj.Geometry(1, 3, 5, 1, 2), 4
    sv = _FakeState(n, geom, L)
    rng = np.random.default_rng(1)
    for special in (True, False, False, True):
        ops_ = _param_circuit(n, rng, special=special)
        prog, _ = program.get_program(sv, ops_, 1)
        ref, _ = o_sim.get_final_state(qb.QuantumScript(ops_, [qb.state()]))
        assert np.max(np.abs(_run_program(prog, n) - ref.reshape(-1))) < 1e-12


def test_canon_batch_matches_scalar_forms():
    from pennylane_b200.program import canon_batch

    rng = np.random.default_rng(3)
    us, forms = [], []
    for _ in range(300):
        kind = rng.integers(5)
        th, ph = rng.uniform(0, 6.3, size=2)
        if rng.random() < 0.15:
            th = float(rng.choice([0.0, np.pi, 2 * np.pi]))
        ry, rx, rz = (np.asarray(g(a, wires=0).matrix()) for g, a in ((q.RY, th), (q.RX, th), (q.RZ, ph)))
        u = [ry, rx, rz @ ry, ry @ rz, _random_unitary(rng)][kind]
        c = sj.canon_1q(u)
        us.append(u)
        forms.append((c.kern, c.dl, c.dr))
    U = np.array(us)
    kern = np.array([f[0] for f in forms]); dl = np.array([f[1] for f in forms]); dr = np.array([f[2] for f in forms])
    t, sinp, r, l, s, ok = canon_batch(U, kern, dl, dr)
    assert np.all(ok)
    for k in range(len(us)):
        c = sj.Canon(int(kern[k]), bool(sinp[k]), float(t[k]), complex(r[k]), complex(l[k]), complex(s[k]))
        assert np.max(np.abs(c.matrix() - us[k])) < 1e-13
        assert (not c.dl or dl[k]) and (not c.dr or dr[k])
    # a general block does not fit the bare kernel
    _, _, _, _, _, ok = canon_batch(np.array([_random_unitary(rng)]), np.array([0]), np.array([False]), np.array([False]))
    assert not ok[0]


def test_small_registers_are_promoted_on_the_second_execution():
    """Below the size threshold a structure runs through the interpreter the first time it is
    seen and is compiled when it comes back (a training loop)."""
    from pennylane_b200 import program

    program.clear()
    n, geom, L = 9, sj.Geometry(1, 3, 5, 1, 2), 4
    sv = _FakeState(n, geom, L, above_threshold=False)
    rng = np.random.default_rng(2)
    prog, hit = program.get_program(sv, _param_circuit(n, rng), 1)
    assert prog is None and not hit
    prog, hit = program.get_program(sv, _param_circuit(n, rng), 1)
    assert prog is not None and not hit
    prog, hit = program.get_program(sv, _param_circuit(n, rng), 1)
    assert prog is not None and hit


def test_adjoint_program_rebinds_parameters():
    """The cached reverse program of one tape, rebound to the parameter values of another tape of
    the same structure, gives that tape's Jacobian (emulated; conj-transposed operator matrices
    multiplied along the merged blocks, generator coefficients corrected by the block scalars)."""
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from oracle.apply_operation import apply_operation as o_apply
    from pennylane_b200 import program
    from pennylane_b200.adjoint import _accumulate_slot_sums, _fused_reverse_program

    n, RB, TB, L = 9, 3, 5, 4
    geom = sj.Geometry(1, RB, TB, 2, 1)

    def tape_for(seed):
        rng = np.random.default_rng(seed)
        ops_ = []
        for l in range(3):
            for w in range(n):
                ops_ += [q.RY(rng.uniform(0, 6), wires=w), q.RZ(rng.uniform(0, 6), wires=w)]
            ops_ += [q.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
            ops_ += [q.RX(rng.uniform(0, 6), wires=1), q.IsingXY(rng.uniform(0, 6), wires=[2, 5]),
                     q.PhaseShift(rng.uniform(0, 6), wires=4), q.CZ(wires=[0, 3]),
                     q.Hadamard(wires=6)]
        obs = q.PauliZ(wires=0) @ q.PauliY(wires=3)
        return qb.QuantumScript(ops_, [qb.expval(obs)]), obs

    tape_a, _ = tape_for(1)
    prims, filled, trainable = _fused_reverse_program(tape_a, n, RB, 1)
    segs = cc.pack_segments(cc.merge_blocks(prims, 1, fold_cx=False), n, T=geom.T, L=L, max_ops=64)
    plans = [sj.plan_segment(s, geom, L) for s in segs]
    prog = program.FusedProgram(segs, plans, len(tape_a.operations))
    assert len(prog.blocks) > 0
    for seed in (2, 3):
        tape_b, obs = tape_for(seed)
        sweep_ops = list(reversed(tape_b.operations))
        tabs = prog.bind(sweep_ops, lambda w: n - 1 - int(w), False, adjoint=True)
        state, _ = o_sim.get_final_state(tape_b)
        ref = np.array(o_adj.adjoint_jacobian(tape_b, state), dtype=float)
        ket = state.reshape(-1).copy()
        bra = 2.0 * o_apply(obs, state).reshape(-1)
        raw, gather = [], []
        for plan, tab in zip(plans, tabs):
            ket, bra, sums = run_plan(plan, tab, ket, n, bra=bra)
            for slot, param in enumerate(plan.slot_params):
                gather.append((len(raw) + slot, 0, param))
            raw += list(-sums)
        jac = _accumulate_slot_sums(np.array(raw), gather, len(trainable), 1)[:, 0]
        assert np.max(np.abs(jac - ref)) < 1e-12
