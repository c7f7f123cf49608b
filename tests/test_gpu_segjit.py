"""GPU parity of the structure-specialised segment kernels (csrc/segk.cuh through
pennylane_b200/segjit.py) against the oracle: forward circuits in both precisions, broadcast
blocks, sharded-style external predicates (base_hi)."""
import numpy as np
import pytest

import pennylane_b200 as qb
from pennylane_b200 import ops as q

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _force_jit(monkeypatch):
    monkeypatch.setenv("B200Q_JIT", "1")


def _state_oracle(ops_, n):
    from oracle import simulate as o_sim

    ref, _ = o_sim.get_final_state(qb.QuantumScript(ops_, [qb.state()]))
    return np.asarray(ref).reshape(-1)


@pytest.mark.parametrize("dtype,tol,level,n,seed", [
    (np.complex128, 1e-12, 1, 13, 1), (np.complex128, 1e-12, 2, 14, 2), (np.complex128, 1e-12, 1, 16, 3),
    (np.complex64, 1e-5, 1, 13, 1), (np.complex64, 1e-5, 2, 15, 4)])
def test_random_circuits_match_oracle(dtype, tol, level, n, seed):
    from pennylane_b200 import segjit
    from pennylane_b200.statevector import StateVector
    from test_segjit import _circuit

    ops_ = _circuit(n, 120, seed=seed)
    ref = _state_oracle(ops_, n)
    sv = StateVector(n, dtype=dtype)
    assert sv.jit_enabled(1)
    before = segjit.stats()
    sv.apply_operations_fused(ops_, level=level)
    got = sv.to_numpy().reshape(-1)
    after = segjit.stats()
    assert after["kernels"] > 0 and (after["compiled"] + after["disk_hits"] + after["mem_hits"]
                                     > before["compiled"] + before["disk_hits"] + before["mem_hits"])
    assert np.max(np.abs(got - ref)) < tol            # tolerance: 1e-12 (c128) / 1e-5 (c64)


@pytest.mark.parametrize("L", [4, 5])
def test_hea_matches_oracle_and_interpreter(L, monkeypatch):
    """The benchmark circuit (RY, RZ, CNOT ring) at 16 qubits: specialised kernels vs the oracle
    (1e-12) and vs the record interpreter."""
    import bench
    from pennylane_b200.statevector import StateVector

    monkeypatch.setenv("B200Q_TILE_L", str(L))
    n = 16
    ops_ = bench.hea_ops(n, layers=4)
    ref = _state_oracle(ops_, n)
    sv = StateVector(n)
    sv.apply_operations_fused(ops_, level=1)
    got = sv.to_numpy().reshape(-1)
    assert np.max(np.abs(got - ref)) < 1e-12
    monkeypatch.setenv("B200Q_JIT", "0")
    sv2 = StateVector(n)
    sv2.apply_operations_fused(ops_, level=1)
    assert np.max(np.abs(sv2.to_numpy().reshape(-1) - got)) < 1e-13


def test_broadcast_blocks():
    """Parameter broadcasting: one coefficient table per batch element."""
    from pennylane_b200.statevector import StateVector

    n, B = 13, 3
    rng = np.random.default_rng(7)
    ops_ = []
    for l in range(3):
        for w in range(n):
            ops_.append(q.RY(rng.uniform(0, 6, size=B), wires=w))
            ops_.append(q.RZ(rng.uniform(0, 6), wires=w))
        ops_ += [q.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    sv = StateVector(n)
    sv.apply_operations_fused(ops_, level=1)
    got = sv.to_numpy().reshape(B, -1)
    for b in range(B):
        ops_b = [type(o)(np.asarray(o.data[0])[b], wires=o.wires) if getattr(o, "batch_size", None) else o
                 for o in ops_]
        assert np.max(np.abs(got[b] - _state_oracle(ops_b, n))) < 1e-12


def test_external_predicates_with_rank_bits():
    """base_hi: controls / parities on bits above the local state (a sharded rank's id)."""
    from pennylane_b200 import compiler as cc
    from pennylane_b200.statevector import StateVector

    n_loc, g = 13, 2
    n = n_loc + g
    rng = np.random.default_rng(11)
    ops_ = []
    for _ in range(40):
        a = int(rng.integers(g, n))
        c = int(rng.integers(0, g))                    # wire on a rank bit: control / diagonal only
        th = rng.uniform(0, 6)
        ops_ += [q.RY(th, wires=a), q.CNOT(wires=[c, a]), q.IsingZZ(th, wires=[c, a]),
                 q.CRY(th, wires=[c, a]), q.RZ(th, wires=c)]
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    from oracle.apply_operation import apply_operation as o_apply
    ref = psi0.reshape((2,) * n)
    for o in ops_:
        ref = o_apply(o, ref)
    ref = ref.reshape(1 << g, -1)
    for rank in range(1 << g):
        sv = StateVector(n_loc)
        sv.set_state(psi0.reshape(1 << g, -1)[rank])
        segs = cc.compile_ops(ops_, n_loc, bit_of=lambda w: n - 1 - int(w), level=1,
                              T=sv.default_tile()[0], L=sv.default_tile()[1], fold_cx=False)
        for seg in segs:
            assert seg.tile_bits is not None
            sv.run_segment(seg, base_hi=rank << n_loc)
        assert np.max(np.abs(sv.to_numpy().reshape(-1) - ref[rank])) < 1e-12


@pytest.mark.parametrize("dtype,tol", [(np.complex128, 1e-12), (np.complex64, 2e-5)])
@pytest.mark.parametrize("n,seed,level", [(13, 1, 1), (14, 2, 1), (13, 3, 0)])
def test_adjoint_jacobian_matches_oracle(dtype, tol, n, seed, level):
    """Fused reverse sweep through the specialised kernels (two vectors per thread, generator
    inner products in the same pass) against the oracle's adjoint_jacobian."""
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from pennylane_b200 import adjoint
    from test_compiler import _trainable_circuit

    ops_ = _trainable_circuit(n, 90, seed=seed)
    obs = [q.PauliZ(wires=0) @ q.PauliX(wires=2), q.PauliY(wires=n - 1)]
    tape = qb.QuantumScript(ops_, [qb.expval(o) for o in obs])
    st, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, st), dtype=float)
    jac = np.array(adjoint.adjoint_jacobian(tape, dtype=dtype, fusion=max(level, 1) if level else 1), dtype=float)
    assert jac.shape == ref.shape
    assert np.max(np.abs(jac - ref)) < tol          # 1e-12 (c128) / 2e-5 (c64)


def test_adjoint_of_the_ansatz_and_split_parameters(monkeypatch):
    """The benchmark tape at 16 qubits, and a small record budget so that the generator terms of
    some parameters straddle segments (the slot sums must be added)."""
    import bench
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from pennylane_b200 import adjoint

    tape = bench.hea_tape(16, 3)
    st, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, st), dtype=float)
    jac = np.array(adjoint.adjoint_jacobian(tape, fusion=1), dtype=float)
    assert np.max(np.abs(jac - ref)) < 1e-12
    n = 13
    rng = np.random.default_rng(5)
    ops_ = []
    for _ in range(20):
        a, b = (int(x) for x in rng.permutation(n)[:2])
        ops_ += [q.RY(rng.uniform(0, 6), wires=a), q.IsingXY(rng.uniform(0, 6), wires=[a, b]),
                 q.PhaseShift(rng.uniform(0, 6), wires=b)]
    tape = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0) @ q.PauliY(wires=3))])
    st, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, st), dtype=float)
    monkeypatch.setattr(adjoint, "MAX_SEGMENT_OPS", 3)
    jac = np.array(adjoint.adjoint_jacobian(tape, fusion=1), dtype=float)
    assert np.max(np.abs(jac - ref)) < 1e-12
    monkeypatch.setenv("B200Q_JIT", "0")             # same split through the record interpreter
    jac = np.array(adjoint.adjoint_jacobian(tape, fusion=1), dtype=float)
    assert np.max(np.abs(jac - ref)) < 1e-12


@pytest.mark.parametrize("n,pb", [(16, 2), (17, 3), (15, 1)])
def test_partial_launches_cover_the_state(n, pb):
    """``fix_mask`` / ``fix_val`` of b200q_seg_launch: running every segment of a circuit piece by
    piece (2**pb partial launches over the tiles whose free index bits equal p) gives the state
    of the whole launches, bit for bit, and the oracle's to 1e-12; the pieces of one segment
    touch disjoint amplitudes (a piece alone leaves the others untouched)."""
    import bench
    from pennylane_b200.sharded import _free_bit_window
    from pennylane_b200.statevector import StateVector

    ops_ = bench.hea_ops(n, layers=3)
    ref = _state_oracle(ops_, n)
    sv = StateVector(n)
    segs = sv.compile_fused(ops_, level=1)
    sv.prepare_segments(segs)
    whole = StateVector(n)
    pieces = 0
    for seg in segs:
        whole.run_segment(seg)
        busy = 0
        for b in seg.tile_bits or range(n):
            busy |= 1 << b
        win = _free_bit_window(busy, n, pb, 0) if sv.segment_partial_ok(seg) else None
        if win is None:
            sv.run_segment(seg)
            continue
        lo, k = win
        mask = ((1 << k) - 1) << lo
        before = sv.to_numpy().reshape(-1).copy()
        sv.run_segment(seg, 0, mask, 0)                       # piece 0 alone
        mid = sv.to_numpy().reshape(-1)
        idx = np.arange(1 << n)
        untouched = (idx & mask) != 0
        assert np.array_equal(mid[untouched], before[untouched])
        for p in range(1, 1 << k):
            sv.run_segment(seg, 0, mask, p << lo)
        pieces += 1
    assert pieces > 0
    got = sv.to_numpy().reshape(-1)
    assert np.array_equal(got, whole.to_numpy().reshape(-1))
    assert np.max(np.abs(got - ref)) < 1e-12                  # tolerance: 1e-12 (complex128)


def test_remap_copy_pack_unpack():
    """b200q_remap_copy: strided piece of a slab -> dense staging (pack) and back (unpack)."""
    import ctypes as C

    import torch

    from pennylane_b200._lib import check, load

    lib = load()
    dev = torch.device("cuda:0")
    src = torch.arange(1 << 16, dtype=torch.float64, device=dev)
    for run, pitch, count, off in [(64, 256, 100, 128), (512, 512, 7, 0), (1, 8, 1000, 3), (4096, 16384, 3, 4096)]:
        stage = torch.zeros(run * count, dtype=torch.float64, device=dev)
        stream = torch.cuda.current_stream().cuda_stream
        check(lib.b200q_remap_copy(C.c_void_p(stage.data_ptr()), run * 8, C.c_void_p(src.data_ptr() + off * 8),
                                   pitch * 8, run * 8, count, C.c_void_p(stream)))
        want = torch.as_strided(src, (count, run), (pitch, 1), off).contiguous().reshape(-1)
        assert torch.equal(stage, want)
        dst = torch.full((1 << 16,), -1.0, dtype=torch.float64, device=dev)
        check(lib.b200q_remap_copy(C.c_void_p(dst.data_ptr() + off * 8), pitch * 8, C.c_void_p(stage.data_ptr()),
                                   run * 8, run * 8, count, C.c_void_p(stream)))
        view = torch.as_strided(dst, (count, run), (pitch, 1), off)
        assert torch.equal(view.contiguous().reshape(-1), want)
        touched = torch.zeros(1 << 16, dtype=torch.bool, device=dev)
        torch.as_strided(touched, (count, run), (pitch, 1), off).fill_(True)
        assert torch.all(dst[~touched] == -1.0)
    # the kernel versions of the unpack copy (TMA bulk copies / registers), pitched and contiguous
    big = torch.arange(1 << 22, dtype=torch.float64, device=dev)
    for mode in (0, 1):
        for run, pitch, count, off in [(2048, 8192, 100, 4096), (1 << 14, 1 << 14, 5, 0), (8, 64, 1000, 8),
                                       (1 << 12, 1 << 15, 64, 1 << 12)]:
            stage = big[: run * count].clone()
            dst = torch.full((1 << 22,), -1.0, dtype=torch.float64, device=dev)
            check(lib.b200q_remap_unpack(C.c_void_p(dst.data_ptr() + off * 8), pitch * 8, C.c_void_p(stage.data_ptr()),
                                         run * 8, run * 8, count, mode, 0, C.c_void_p(stream)))
            view = torch.as_strided(dst, (count, run), (pitch, 1), off)
            assert torch.equal(view.contiguous().reshape(-1), stage), (mode, run, pitch, count)
            touched = torch.zeros(1 << 22, dtype=torch.bool, device=dev)
            torch.as_strided(touched, (count, run), (pitch, 1), off).fill_(True)
            assert torch.all(dst[~touched] == -1.0)
