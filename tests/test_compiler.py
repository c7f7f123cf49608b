"""CPU: the host fusion pass (lowering -> block merging -> segment packing) preserves the
circuit.  A small numpy interpreter of the tile primitives (test-only) replays the compiled
segments and is compared with the oracle's gate-by-gate result."""
import numpy as np
import pytest

from conftest import random_state
from oracle.apply_operation import apply_operation
from pennylane_b200 import ops as q
from pennylane_b200.compiler import (CX, DENSE1, DENSE2, DIAG, GENERIC, PARITY, SWAP, TileOp,
                                     compile_ops, encode_segment)


def _apply_prim(p, psi, n):
    """psi: flat array indexed by the integer whose bit b is `bit position b`."""
    N = 1 << n
    idx = np.arange(N)
    sel = np.ones(N, dtype=bool)
    for b, v in p.ctrl.items():
        sel &= ((idx >> b) & 1) == v
    out = psi.copy()
    if p.kind == PARITY:
        par = np.zeros(N, dtype=int)
        for b in p.other:
            par ^= (idx >> b) & 1
        ph = np.where(par == 1, p.mat[1], p.mat[0])
        out[sel] = psi[sel] * ph[sel]
    elif p.kind == DIAG:
        k = len(p.other)
        t = np.zeros(N, dtype=int)
        for j, b in enumerate(p.other):
            t |= ((idx >> b) & 1) << (k - 1 - j)
        out[sel] = psi[sel] * np.asarray(p.mat)[t[sel]]
    elif p.kind in (CX, DENSE1):
        b = p.targets[0]
        variants = [(sel, np.array([[0, 1], [1, 0]], dtype=complex) if p.kind == CX else p.mat)]
        if p.kind == DENSE1 and p.mat0 is not None:
            variants.append((~sel, p.mat0))          # controlled-select block
        for s_, m in variants:
            i0 = idx[s_ & (((idx >> b) & 1) == 0)]
            i1 = i0 | (1 << b)
            out[i0] = m[0, 0] * psi[i0] + m[0, 1] * psi[i1]
            out[i1] = m[1, 0] * psi[i0] + m[1, 1] * psi[i1]
    elif p.kind in (DENSE2, SWAP):
        m = p.mat if p.kind == DENSE2 else np.eye(4)[[0, 2, 1, 3]].astype(complex)
        b0, b1 = p.targets
        base = idx[sel & (((idx >> b0) & 1) == 0) & (((idx >> b1) & 1) == 0)]
        ii = [base, base | (1 << b1), base | (1 << b0), base | (1 << b0) | (1 << b1)]
        x = [psi[i] for i in ii]
        for r in range(4):
            out[ii[r]] = sum(m[r, c] * x[c] for c in range(4))
    else:
        raise AssertionError(p.kind)
    return out


def _replay(segs, state, n):
    psi = state.reshape(-1).copy()
    for seg in segs:
        for p in seg.prims:
            if seg.tile_bits is not None:
                assert set(p.targets) <= set(seg.tile_bits), "target outside the tile"
            if p.kind == GENERIC:
                psi = apply_operation(p.op, psi.reshape((2,) * n)).reshape(-1)
            else:
                psi = _apply_prim(p, psi, n)
    return psi.reshape((2,) * n)


def _random_circuit(n, depth, seed):
    rng = np.random.default_rng(seed)
    ops_ = []
    for _ in range(depth):
        w = [int(x) for x in rng.permutation(n)]
        a, b, c = w[0], w[1], w[2]
        th = rng.uniform(0, 6)
        choices = [
            q.RX(th, wires=a), q.RY(th, wires=a), q.RZ(th, wires=a), q.Hadamard(wires=a),
            q.PauliX(wires=a), q.PauliY(wires=a), q.PauliZ(wires=a), q.S(wires=a), q.T(wires=a),
            q.SX(wires=a), q.PhaseShift(th, wires=a), q.Rot(th, 0.3, -th, wires=a),
            q.GlobalPhase(th, wires=a), q.CNOT(wires=[a, b]), q.CZ(wires=[a, b]),
            q.CY(wires=[a, b]), q.SWAP(wires=[a, b]), q.ISWAP(wires=[a, b]),
            q.CRX(th, wires=[a, b]), q.CRZ(th, wires=[a, b]), q.CRot(th, .1, .2, wires=[a, b]),
            q.ControlledPhaseShift(th, wires=[a, b]), q.IsingXX(th, wires=[a, b]),
            q.IsingZZ(th, wires=[a, b]), q.IsingXY(th, wires=[a, b]),
            q.PauliRot(th, "XY", wires=[a, b]), q.PauliRot(th, "ZIZ", wires=[a, b, c]),
            q.PauliRot(th, "XYZ", wires=[a, b, c]), q.MultiRZ(th, wires=[a, b, c]),
            q.Toffoli(wires=[a, b, c]), q.CSWAP(wires=[a, b, c]), q.CCZ(wires=[a, b, c]),
            q.MultiControlledX(wires=[a, b, c], control_values=[0, 1]),
            q.ctrl(q.RY(th, wires=a), [b, c], [1, 0]), q.ctrl(q.IsingXX(th, wires=[a, b]), c),
            q.adjoint(q.S(wires=a)), q.DiagonalQubitUnitary(np.exp(1j * rng.normal(size=8)), wires=[a, b, c]),
            q.QubitUnitary(np.linalg.qr(rng.normal(size=(8, 8)) + 1j * rng.normal(size=(8, 8)))[0], wires=[a, b, c]),
        ]
        ops_.append(choices[int(rng.integers(len(choices)))])
    return ops_


@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("n,T,L", [(6, 4, 2), (7, 5, 3), (8, 12, 5), (9, 6, 0)])
def test_compiled_segments_reproduce_circuit(level, n, T, L):
    ops_ = _random_circuit(n, 120, seed=n * 10 + level)
    state = random_state(n, seed=n)
    ref = state
    for op in ops_:
        ref = apply_operation(op, ref)
    segs = compile_ops(ops_, n, level=level, T=T, L=L)
    assert sum(s.ngates for s in segs) == sum(1 for o in ops_ if o.name != "Identity")
    for s in segs:
        if s.tile_bits is not None:
            assert len(s.tile_bits) == min(T, n) and s.tile_bits == sorted(s.tile_bits)
            assert s.tile_bits[:min(L, n)] == list(range(min(L, n)))
    got = _replay(segs, state, n)
    assert np.max(np.abs(got - ref)) < 1e-12


def test_hea_packing_and_encoding():
    import bench

    n = 30
    ops_ = bench.hea_ops(n)
    segs = compile_ops(ops_, n, level=1, T=12, L=5)
    assert sum(s.ngates for s in segs) == 720 and len(segs) <= 40
    for s in segs:
        arr, table = encode_segment(s)
        assert len(arr) >= len(s.prims) and table.dtype == np.complex128
        assert isinstance(arr[0], TileOp)
        for o in arr:
            assert 0 <= o.t0 < 12 and 0 <= o.t1 < 12 and o.mat_off + 2 <= table.size + 2


def test_merging_reduces_primitives():
    ops_ = [q.RZ(.1, wires=0), q.RY(.2, wires=0), q.RZ(.3, wires=0), q.Hadamard(wires=1),
            q.CNOT(wires=[0, 1]), q.T(wires=1), q.S(wires=1)]
    l0 = compile_ops(ops_, 2, level=0)
    l1 = compile_ops(ops_, 2, level=1)
    assert sum(len(s.prims) for s in l0) == 7 and sum(len(s.prims) for s in l1) == 2
    assert sum(s.ngates for s in l1) == 7


# ---- register-tiled kernel: round scheduling + record encoding (CPU, via tests/rt_emulator) ----
@pytest.mark.parametrize("level", [0, 1, 2])
@pytest.mark.parametrize("n,T,L,RB,sww", [(9, 7, 3, 3, 3), (10, 8, 5, 3, 3), (12, 12, 5, 4, 3),
                                          (13, 12, 4, 4, 3), (13, 13, 5, 5, 4), (13, 12, 5, 3, 3)])
def test_round_schedule_reproduces_circuit(level, n, T, L, RB, sww):
    from pennylane_b200.compiler import RT_ROUND, encode_rt_segment, schedule_rounds
    from rt_emulator import run_records

    ops_ = _random_circuit(n, 100, seed=n * 7 + level + RB)
    state = random_state(n, seed=n + 1)
    ref = state
    for op in ops_:
        ref = apply_operation(op, ref)
    segs = compile_ops(ops_, n, level=level, T=T, L=L)
    psi = state.reshape(-1).copy()
    from pennylane_b200.compiler import _IO_LANES
    lanes = min(_IO_LANES, T - RB)
    for seg in segs:
        if seg.tile_bits is None:
            p = seg.prims[0]
            if p.kind == GENERIC:
                psi = apply_operation(p.op, psi.reshape((2,) * n)).reshape(-1)
            else:
                psi = _apply_prim(p, psi, n)
            continue
        rounds = schedule_rounds(seg.prims, seg.tile_bits, RB, sww)
        # the last (store) round keeps tile positions 0..lanes-1 on the lanes; the first round is
        # read from the landing buffer: only positions 0..sww-1 must stay on the lowest lane bits
        assert rounds[-1].tpos[:lanes] == list(range(lanes))
        assert all(p >= lanes for p in rounds[-1].rpos)
        low = min(sww, T - RB)
        assert rounds[0].tpos[:low] == list(range(low))
        assert all(p >= low for p in rounds[0].rpos)
        for r in rounds:
            assert sorted(r.rpos + r.tpos) == list(range(T))
        arr, table, nrec = encode_rt_segment(seg, RB, sww, rounds=rounds)
        assert arr[0].kind == RT_ROUND and nrec == len(arr)
        psi = run_records(psi, n, seg.tile_bits, arr, table, RB)
    assert np.max(np.abs(psi.reshape((2,) * n) - ref)) < 1e-12


def test_round_schedule_hea_round_count():
    """The 30-qubit ansatz needs few shared-memory transposes per segment."""
    import bench
    from pennylane_b200.compiler import schedule_rounds

    segs = compile_ops(bench.hea_ops(30), 30, level=1, T=12, L=5)
    nrounds = [len(schedule_rounds(s.prims, s.tile_bits, 4)) for s in segs]
    assert max(nrounds) <= 6 and sum(nrounds) <= 4 * len(segs)


# ---- fused adjoint reverse sweep: program construction + GEN records (CPU, emulated) -----------
def _trainable_circuit(n, depth, seed):
    rng = np.random.default_rng(seed)
    ops_ = []
    for _ in range(depth):
        w = [int(x) for x in rng.permutation(n)]
        a, b, c = w[0], w[1], w[2]
        th = rng.uniform(0, 6)
        choices = [q.RX(th, wires=a), q.RY(th, wires=a), q.RZ(th, wires=a), q.PhaseShift(th, wires=a),
                   q.IsingXX(th, wires=[a, b]), q.IsingYY(th, wires=[a, b]), q.IsingZZ(th, wires=[a, b]),
                   q.PauliRot(th, "XY", wires=[a, b]), q.PauliRot(th, "ZZ", wires=[a, b]),
                   q.MultiRZ(th, wires=[a, b, c]), q.CNOT(wires=[a, b]), q.Hadamard(wires=a),
                   q.CZ(wires=[a, b]), q.Toffoli(wires=[a, b, c]), q.S(wires=a), q.SWAP(wires=[a, b])]
        ops_.append(choices[int(rng.integers(len(choices)))])
    return ops_


@pytest.mark.parametrize("level", [0, 1])
@pytest.mark.parametrize("n,T,L,RB", [(8, 7, 3, 3), (9, 8, 4, 3), (10, 9, 5, 4)])
def test_fused_reverse_sweep_program(level, n, T, L, RB):
    import pennylane_b200 as qb
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from pennylane_b200.adjoint import _fused_reverse_program
    from pennylane_b200.compiler import GEN, encode_rt_segment, merge_blocks, pack_segments
    from rt_emulator import run_records

    ops_ = _trainable_circuit(n, 60, seed=n + level)
    obs = q.PauliZ(wires=0) @ q.PauliX(wires=2)
    tape = qb.QuantumScript(ops_, [qb.expval(obs)])
    state, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, state), dtype=float)
    prog = _fused_reverse_program(tape, n, RB, level)
    assert prog is not None
    prims, filled, trainable = prog
    segs = pack_segments(merge_blocks(prims, level), n, T=T, L=L, max_ops=64)
    ket = state.reshape(-1).copy()
    from oracle.apply_operation import apply_operation as o_apply
    bra = 2.0 * o_apply(obs, state).reshape(-1)
    jac = np.zeros(len(trainable))
    for seg in segs:
        assert seg.tile_bits is not None
        local = {}
        for p in seg.prims:
            if p.kind == GEN:
                p.slot = local.setdefault(p.param, len(local))
        arr, table, nrec = encode_rt_segment(seg, RB)
        ket, bra, sums = run_records(ket, n, seg.tile_bits, arr, table, RB, bra=bra, nslots=len(local))
        for param, slot in local.items():
            jac[param] += -sums[slot]
    assert sorted(filled) == list(range(len(trainable)))
    assert np.max(np.abs(jac - ref)) < 1e-12


def test_reverse_sweep_keeps_one_dense_record_per_wire_and_layer():
    """Generator terms are hoisted through a pending CNOT (Z_t -> Z_c Z_t) or deferred behind the
    block they commute with (compiler._place_generator): the reverse sweep of [RY, RZ] + CNOT ring
    has ONE dense record per (wire, layer), not one per rotation — and still the oracle's
    Jacobian through the record emulator."""
    import pennylane_b200 as qb
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    from oracle.apply_operation import apply_operation as o_apply
    from pennylane_b200.adjoint import _fused_reverse_program
    from pennylane_b200.compiler import GEN, encode_rt_segment, merge_blocks, pack_segments
    from rt_emulator import run_records

    n, layers, T, L, RB = 10, 3, 9, 5, 3
    rng = np.random.default_rng(9)
    ops_ = []
    for _ in range(layers):
        for w in range(n):
            ops_ += [q.RY(rng.uniform(0, 6), wires=w), q.RZ(rng.uniform(0, 6), wires=w)]
        ops_ += [q.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    obs = q.PauliZ(wires=0) @ q.PauliY(wires=3)
    tape = qb.QuantumScript(ops_, [qb.expval(obs)])
    prims, filled, trainable = _fused_reverse_program(tape, n, RB, 1)
    merged = merge_blocks(prims, 1)
    assert sum(p.kind == DENSE1 for p in merged) == n * layers
    assert sum(p.kind == GEN for p in merged) == 2 * n * layers
    # hoisted Z terms picked up the CNOT's control: two Z factors, no X part
    assert any(p.kind == GEN and len(p.zbits) == 2 and not p.targets for p in merged)
    state, _ = o_sim.get_final_state(tape)
    ref = np.array(o_adj.adjoint_jacobian(tape, state), dtype=float)
    ket = state.reshape(-1).copy()
    bra = 2.0 * o_apply(obs, state).reshape(-1)
    jac = np.zeros(len(trainable))
    for seg in pack_segments(merged, n, T=T, L=L, max_ops=64):
        local = {}
        for p in seg.prims:
            if p.kind == GEN:
                p.slot = local.setdefault(p.param, len(local))
        arr, table, nrec = encode_rt_segment(seg, RB)
        ket, bra, sums = run_records(ket, n, seg.tile_bits, arr, table, RB, bra=bra, nslots=len(local))
        for param, slot in local.items():
            jac[param] += -sums[slot]
    assert sorted(filled) == list(range(len(trainable)))
    assert np.max(np.abs(jac - ref)) < 1e-12


from pennylane_b200 import compiler as cc  # noqa: E402


def test_trim_rounds_keeps_every_primitive_and_saves_rounds():
    """pack_segments(trim_rounds=True): trailing single-qubit blocks are handed to later segments
    when that saves a round; no primitive is lost or duplicated, the segment count stays, the
    total number of rounds does not grow, and the circuit still reproduces the oracle's state
    (through the plan emulator)."""
    import bench
    from pennylane_b200 import segjit as sj

    n = 30
    geom = sj.default_geometry(1, 1)
    ops_ = bench.hea_ops(n, 8)
    prims = cc.merge_blocks(cc.lower_all(ops_, lambda w: n - 1 - int(w), False), 1, False)
    out = {}
    for trim in (False, True):
        segs = cc.pack_segments(list(prims), n, T=geom.T, L=5, RB=geom.RB, sww=geom.sww, trim_rounds=trim)
        ids = [id(p) for s in segs for p in s.prims]
        assert sorted(ids) == sorted(id(p) for p in prims)
        rounds = [len(s.rounds or cc.schedule_rounds(s.prims, s.tile_bits, geom.RB, geom.sww)) for s in segs]
        out[trim] = (len(segs), sum(rounds), max(rounds))
    assert out[True][0] == out[False][0]
    assert out[True][1] < out[False][1] and out[True][2] < out[False][2]


def test_store_lanes_follow_the_contiguous_low_run():
    """compiler.io_lanes: only the tile's contiguous low run stays on the lanes in the store round
    (a position above the run is not contiguous in memory); with L = 4 tiles an 8-target segment
    then fits two rounds."""
    assert cc.low_run([0, 1, 2, 3, 7, 9]) == 4 and cc.low_run([0, 1, 2, 3, 4, 5]) == 6 and cc.low_run([1, 2]) == 0
    assert cc.io_lanes([0, 1, 2, 3, 10, 11, 12, 13, 14, 15, 16, 17], 4) == 4
    assert cc.io_lanes([0, 1, 2, 3, 4, 11, 12, 13, 14, 15, 16, 17], 4) == 5
    assert cc.io_lanes(list(range(12)), 4) == min(cc._IO_LANES, 8)
    import bench
    from pennylane_b200 import segjit as sj

    geom = sj.default_geometry(1, 1)
    n = 30
    segs = cc.compile_ops(bench.hea_ops(n, 8), n, level=1, T=geom.T, L=4, fold_cx=False, RB=geom.RB, sww=geom.sww)
    rounds = [len(s.rounds or cc.schedule_rounds(s.prims, s.tile_bits, geom.RB, geom.sww)) for s in segs]
    assert len(segs) <= 27 and sum(rounds) <= 70
    for s in segs:
        last = (s.rounds or cc.schedule_rounds(s.prims, s.tile_bits, geom.RB, geom.sww))[-1]
        lanes = cc.io_lanes(s.tile_bits, geom.RB)
        assert all(r >= lanes for r in last.rpos) and last.tpos[:lanes] == list(range(lanes))
