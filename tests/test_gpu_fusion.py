"""GPU parity: the fused path (host fusion pass + shared-memory tile kernel) against the oracle's
gate-by-gate result, for every primitive kind, both precisions, several tile geometries and
batched states."""
import numpy as np
import pytest

from conftest import TOL, random_state

pytestmark = pytest.mark.gpu


def _run_fused(ops_, state, level, T, L, dtype=np.complex128, batched=False):
    from pennylane_b200 import StateVector

    n = state.ndim - (1 if batched else 0)
    sv = StateVector(n, dtype=dtype)
    sv.set_state(state.astype(dtype))
    nseg = sv.apply_operations_fused(ops_, level=level, T=T, L=L)
    return sv.to_numpy(), nseg


def _oracle(ops_, state, batched=False):
    from oracle.apply_operation import apply_operation

    ref = state
    for op in ops_:
        ref = apply_operation(op, ref, is_state_batched=batched)
    return ref


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("n,T,L", [(5, 12, 5), (8, 6, 3), (10, 7, 0), (13, 12, 5), (14, 10, 6), (16, 12, 4)])
def test_random_circuits_all_primitives(level, n, T, L):
    from test_compiler import _random_circuit

    ops_ = _random_circuit(n, 150, seed=100 + n + level)
    state = random_state(n, seed=n)
    got, nseg = _run_fused(ops_, state, level, T, L)
    ref = _oracle(ops_, state)
    assert np.max(np.abs(got - ref)) < 1e-12
    assert nseg < len(ops_)


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_default_tile_geometry_hea(dtype):
    import bench

    n = 18
    ops_ = bench.hea_ops(n, layers=3)
    from pennylane_b200 import StateVector

    sv = StateVector(n, dtype=dtype)
    nseg = sv.apply_operations_fused(ops_, level=1)
    state0 = np.zeros((2,) * n, dtype=complex)
    state0[(0,) * n] = 1
    ref = _oracle(ops_, state0)
    assert np.max(np.abs(sv.to_numpy() - ref)) < TOL[np.dtype(dtype)]
    assert nseg <= len(ops_) // 8


def test_low_bit_gates_bank_conflict_paths():
    """Gates on tile bits 0..2 take the flipped load order in the kernel."""
    from pennylane_b200 import ops as q

    n = 9
    ops_ = []
    for w in (n - 1, n - 2, n - 3, n - 4):
        ops_ += [q.RY(0.3 + w, wires=w), q.Hadamard(wires=w), q.RX(0.1 * w, wires=w)]
    ops_ += [q.CNOT(wires=[n - 1, n - 2]), q.CNOT(wires=[n - 3, n - 1]), q.SWAP(wires=[n - 1, n - 3]),
             q.IsingXX(0.4, wires=[n - 2, n - 1]), q.CRY(0.5, wires=[n - 1, n - 2]),
             q.Toffoli(wires=[0, n - 1, n - 2])]
    state = random_state(n, seed=2)
    for level in (0, 1, 2):
        got, _ = _run_fused(ops_, state, level, 9, 5)
        assert np.max(np.abs(got - _oracle(ops_, state))) < 1e-12


def test_batched_state_and_broadcast_fallback():
    from pennylane_b200 import ops as q

    n, B = 7, 3
    state = random_state(n, seed=3, batch=B)
    th = np.array([0.1, 0.2, 0.3])
    ops_ = [q.Hadamard(wires=0), q.CNOT(wires=[0, 4]), q.RX(th, wires=2), q.RZ(0.3, wires=2),
            q.IsingZZ(th, wires=[1, 6]), q.T(wires=6), q.CRX(0.4, wires=[5, 3])]
    got, _ = _run_fused(ops_, state, 1, 6, 3, batched=True)
    ref = _oracle(ops_, state, batched=True)
    assert got.shape == ref.shape and np.max(np.abs(got - ref)) < 1e-12


def test_device_results_identical_with_and_without_fusion():
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from test_gpu_device import _sel_tape

    n = 12
    tape = _sel_tape(n, 3, 7)
    vals = [qb.B200Qubit(wires=n, fusion=f).execute(tape) for f in (0, 1, 2)]
    assert abs(vals[0] - vals[1]) < 1e-13 and abs(vals[0] - vals[2]) < 1e-13
    (pt,), cfg = qb.B200Qubit(wires=n).preprocess(tape, qb.ExecutionConfig(gradient_method="adjoint"))
    j0 = np.array(qb.B200Qubit(wires=n, fusion=0).compute_derivatives(pt), dtype=float)
    j1 = np.array(qb.B200Qubit(wires=n, fusion=1).compute_derivatives(pt), dtype=float)
    assert np.max(np.abs(j0 - j1)) < 1e-13
    s0 = qb.B200Qubit(wires=n, seed=3, fusion=0).execute(
        qb.QuantumScript(tape.operations, [qb.sample(wires=range(n))], shots=2000))
    s1 = qb.B200Qubit(wires=n, seed=3, fusion=1).execute(
        qb.QuantumScript(tape.operations, [qb.sample(wires=range(n))], shots=2000))
    assert np.mean(np.any(s0 != s1, axis=1)) < 5e-3     # rounding may move boundary shots
    del q


# ---- register-tiled kernel (b200q_apply_rtile) ----------------------------------------------
@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("dtype,n,L", [(np.complex128, 12, 5), (np.complex128, 14, 5),
                                       (np.complex128, 15, 3), (np.complex128, 16, 6),
                                       (np.complex64, 13, 5), (np.complex64, 15, 4),
                                       (np.complex64, 16, 7)])
def test_rtile_random_circuits(level, dtype, n, L):
    """Every primitive kind through the register-tiled kernel (the tile geometry is the
    kernel's own: T = 12 for complex128, 13 for complex64)."""
    from pennylane_b200 import StateVector
    from test_compiler import _random_circuit

    ops_ = _random_circuit(n, 160, seed=300 + n + level)
    state = random_state(n, seed=n)
    sv = StateVector(n, dtype=dtype)
    T = sv.rt_geometry(1)[0]
    sv.set_state(state.astype(dtype))
    nseg = sv.apply_operations_fused(ops_, level=level, T=T, L=L)
    ref = _oracle(ops_, state)
    assert np.max(np.abs(sv.to_numpy() - ref)) < TOL[np.dtype(dtype)] * (1 if dtype == np.complex128 else 3)
    assert nseg < len(ops_)


def test_rtile_batched_states_share_the_program():
    from pennylane_b200 import StateVector
    from test_compiler import _random_circuit

    n, B = 13, 3
    ops_ = _random_circuit(n, 60, seed=11)
    state = random_state(n, seed=5, batch=B)
    sv = StateVector(n, batch=B)
    sv.set_state(state)
    sv.apply_operations_fused(ops_, level=1)
    ref = _oracle(ops_, state, batched=True)
    assert np.max(np.abs(sv.to_numpy() - ref)) < 1e-12


@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_rtile_broadcast_parameters_are_fused(dtype):
    """Gates with a leading batch axis on their parameters (apply_operation.py:186-197) go through
    the register-tiled kernel as dense blocks with one matrix table per batch element
    (b200q_apply_rtile_bcast), merged with their unbatched neighbours and folded CNOTs."""
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.compiler import GENERIC, compile_ops
    from oracle import simulate as o_sim

    n, B = 14, 3
    rng = np.random.default_rng(44)
    ops_ = []
    for layer in range(3):
        for i in range(n):
            ops_ += [q.RY(rng.uniform(0, 6, B), wires=i), q.RZ(rng.uniform(0, 6, B), wires=i)]
            if i % 3 == 0:
                ops_.append(q.Hadamard(wires=i))
            if i % 4 == 1:
                ops_.append(q.Rot(rng.uniform(0, 6, B), 0.3, rng.uniform(0, 6, B), wires=i))
        ops_ += [q.CNOT(wires=[i, (i + 1 + layer) % n]) for i in range(n)]
        ops_ += [q.RX(rng.uniform(0, 6, B), wires=i) for i in range(0, n, 2)]
        ops_.append(q.IsingZZ(0.4, wires=[1, 5]))
    T = qb.StateVector(n, dtype=dtype).rt_geometry(1)[0]
    segs = compile_ops(ops_, n, level=1, T=T, L=5, batched_ok=True)
    assert all(p.kind != GENERIC for s in segs for p in s.prims) and len(segs) < 20
    heis = q.LinearCombination(
        [1.0] * (3 * (n - 1)),
        [P(wires=i) @ P(wires=i + 1) for i in range(n - 1) for P in (q.PauliX, q.PauliY, q.PauliZ)])
    tape = qb.QuantumScript(ops_, [qb.expval(heis), qb.probs(wires=[0, 7, 13])])
    e, p = qb.B200Qubit(c_dtype=dtype, fusion=1).execute(tape)
    re, rp = o_sim.simulate(tape)
    tol = TOL[np.dtype(dtype)]
    assert e.shape == (B,) and p.shape == (B, 8)
    assert np.max(np.abs(e - re)) < tol * 30 and np.max(np.abs(p - rp)) < tol * 3


def test_rtile_hea_matches_unfused_at_20_qubits():
    import bench
    from pennylane_b200 import StateVector

    n = 20
    ops_ = bench.hea_ops(n, layers=2)
    a = StateVector(n)
    a.apply_operations_fused(ops_, level=1)
    b = StateVector(n)
    for op in ops_:
        b.apply_operation(op)
    assert np.max(np.abs(a.to_numpy() - b.to_numpy())) < 1e-13


def test_differential_fuzz_of_the_fused_paths():
    """tools/fuzz_gpu.py: random circuits (all primitive kinds, broadcast rotations, random L /
    fusion level / precision) and random trainable circuits' Jacobians against the oracle."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "tools", "fuzz_gpu.py"), "16"],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
