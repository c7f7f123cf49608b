"""Sharded statevector: host logic under ``gloo`` with world sizes 2 and 4 on CPU ranks
(SURVEY.md section 8(e)).  The numerical kernels are the oracle's (tests/np_engine.py); what is
checked here is the planner, the operator specialisation by rank bits, the in-place exchange,
the ordered reductions and the distributed sampler — against the oracle's single-array
simulation of the same circuit."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn_name, args, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        res = globals()[fn_name](rank, world, *args)
        q.put((rank, "ok", res))
    except Exception:                                   # noqa: BLE001
        q.put((rank, "err", traceback.format_exc()))
    finally:
        if dist.is_initialized():
            dist.destroy_process_group()


def run_ranks(world, fn_name, *args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, fn_name, args, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    for _ in range(world):
        rank, status, res = q.get(timeout=180)
        assert status == "ok", f"rank {rank} failed:\n{res}"
        out[rank] = res
    for p in procs:
        p.join(timeout=30)
    return [out[r] for r in range(world)]


# ---------------------------------------------------------------------------------------------
# circuits
# ---------------------------------------------------------------------------------------------
def _mixed_circuit(n, seed):
    """Every specialisation branch: controls / diagonals / parities / Pauli rotations on wires
    that start (and later become) rank bits, plus dense gates that force remaps."""
    from pennylane_b200 import ops as q

    rng = np.random.default_rng(seed)
    a = lambda: float(rng.uniform(0, 2 * np.pi))          # noqa: E731
    ops_ = [q.Hadamard(wires=w) for w in range(n)]
    ops_ += [q.RZ(a(), wires=0), q.PauliZ(wires=1), q.S(wires=0), q.T(wires=1),
             q.PhaseShift(a(), wires=0), q.CZ(wires=[0, n - 1]), q.CZ(wires=[0, 1]),
             q.ControlledPhaseShift(a(), wires=[1, 2]), q.CCZ(wires=[0, 1, 3]),
             q.MultiRZ(a(), wires=[0, 2, n - 1]), q.IsingZZ(a(), wires=[0, 1]),
             q.CRZ(a(), wires=[0, 3]), q.CRZ(a(), wires=[3, 0]), q.CRZ(a(), wires=[0, 1]),
             q.GlobalPhase(a()),
             q.PauliRot(a(), "ZXY", wires=[0, 2, 3]), q.PauliRot(a(), "ZZ", wires=[0, 1]),
             q.PauliRot(a(), "IZ", wires=[0, 1]),
             q.DiagonalQubitUnitary(np.exp(1j * rng.uniform(0, 6, 8)), wires=[0, 2, 1]),
             q.CNOT(wires=[0, 2]), q.Toffoli(wires=[0, 1, 3]), q.Toffoli(wires=[0, 2, 3]),
             q.CY(wires=[1, 2]), q.CH(wires=[0, 3]), q.CSWAP(wires=[0, 2, 3]),
             q.CRX(a(), wires=[0, 2]), q.CRY(a(), wires=[1, 3]), q.CRot(a(), a(), a(), wires=[0, 2]),
             q.MultiControlledX(wires=[0, 1, 2, 3], control_values=[1, 0, 1]),
             q.MultiControlledX(wires=[0, 1, 2], control_values=[0, 0]),
             q.Controlled(q.RX(a(), wires=2), [0, 1], [True, False]),
             q.Controlled(q.IsingXX(a(), wires=[2, 3]), [0]),
             q.ControlledQubitUnitary(q.RY.compute_matrix(a()), wires=[1, 0, 3])]
    # dense gates on the rank-bit wires: remaps
    ops_ += [q.RY(a(), wires=0), q.RX(a(), wires=1), q.CNOT(wires=[2, 0]), q.IsingXX(a(), wires=[0, 1]),
             q.SWAP(wires=[0, n - 1]), q.Rot(a(), a(), a(), wires=1), q.CNOT(wires=[n - 1, 1]),
             q.PauliRot(a(), "XYZ", wires=[0, 1, 2]), q.Hadamard(wires=0), q.RZ(a(), wires=0),
             q.CZ(wires=[0, 1]), q.RY(a(), wires=n - 1), q.CNOT(wires=[0, 1]), q.RX(a(), wires=0)]
    return ops_


def _hea(n, layers, seed):
    from pennylane_b200 import ops as q

    par = np.random.default_rng(seed).uniform(0, 2 * np.pi, (layers, n, 2))
    ops_ = []
    for l in range(layers):
        for w in range(n):
            ops_ += [q.RY(par[l, w, 0], wires=w), q.RZ(par[l, w, 1], wires=w)]
        for w in range(n):
            ops_.append(q.CNOT(wires=[w, (w + 1) % n]))
    return ops_


def _oracle_state(n, ops_):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim

    st, _ = o_sim.get_final_state(qb.QuantumScript(ops_, [qb.state()] if hasattr(qb, "state") else []))
    return np.asarray(st).reshape(-1)


def _sharded(n, world, batch=1):
    from np_engine import NumpyEngine
    from pennylane_b200.sharded import ShardedStateVector

    g = world.bit_length() - 1
    return ShardedStateVector(n, dist, engine=NumpyEngine(n - g, batch), stage_bytes=4096)


# ---------------------------------------------------------------------------------------------
# rank functions
# ---------------------------------------------------------------------------------------------
def w_state(rank, world, n, kind, seed):
    ops_ = _mixed_circuit(n, seed) if kind == "mixed" else _hea(n, 3, seed)
    sv = _sharded(n, world)
    sv.apply_operations(ops_)
    got = sv.to_numpy()
    ref = _oracle_state(n, ops_)
    return float(np.max(np.abs(got - ref))), dict(sv.stats), list(sv.phys)


def w_window(rank, world, n, seed, piece_bits, stage_bytes):
    """The overlapped schedule (ShardedStateVector._schedule / _run_window): segments on either
    side of an exchange run piece by piece, the exchange moves one piece at a time."""
    from np_engine import NumpyEngine
    from pennylane_b200.sharded import ShardedStateVector

    os.environ["B200Q_EXCHANGE_PIECE_BITS"] = str(piece_bits)
    os.environ["B200Q_EXCHANGE_MIN_BIT"] = "0"          # tiny shards: any index bit may cut the pieces
    try:
        ops_ = _hea(n, 3, seed) + _mixed_circuit(n, seed + 1)
        g = world.bit_length() - 1
        eng = NumpyEngine(n - g)
        sv = ShardedStateVector(n, dist, engine=eng, stage_bytes=stage_bytes)
        prog = sv.compile(ops_)
        kinds = [e[0] for e in prog["schedule"]]
        sv.run(prog)
        got = sv.to_numpy()
        ref = _oracle_state(n, ops_)
        return (float(np.max(np.abs(got - ref))), kinds.count("window"), kinds.count("exchange"),
                getattr(eng, "partial_runs", 0), dict(sv.stats))
    finally:
        os.environ.pop("B200Q_EXCHANGE_PIECE_BITS", None)
        os.environ.pop("B200Q_EXCHANGE_MIN_BIT", None)


def w_expval(rank, world, n, seed):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q

    ops_ = _hea(n, 2, seed)
    # Heisenberg chain + a word with Y factors on the rank-bit wires
    H = None
    for i in range(n - 1):
        for P in (q.PauliX, q.PauliY, q.PauliZ):
            t = P(wires=i) @ P(wires=i + 1)
            H = t if H is None else H + t
    H = H + 0.7 * (q.PauliY(wires=0) @ q.PauliZ(wires=1) @ q.PauliX(wires=n - 1))
    sv = _sharded(n, world)
    sv.apply_operations(ops_)
    got = sv.expval_pauli_sentence(H.pauli_rep)
    z0 = sv.expval_pauli_sentence(q.PauliZ(wires=0).pauli_rep)
    norm = sv.norm2()
    tape = qb.QuantumScript(ops_, [qb.expval(H), qb.expval(q.PauliZ(wires=0))])
    ref = o_sim.simulate(tape)
    return abs(got - ref[0]), abs(z0 - ref[1]), abs(norm - 1.0), dict(sv.stats)


def w_probs(rank, world, n, seed):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim

    ops_ = _mixed_circuit(n, seed)
    sv = _sharded(n, world)
    sv.apply_operations(ops_)
    errs = []
    for wires in ([0], [1, 0], [n - 1, 0, 2], list(range(n)), [2, 3]):
        ref = o_sim.simulate(qb.QuantumScript(ops_, [qb.probs(wires=wires)]))
        errs.append(float(np.max(np.abs(sv.probs(wires) - ref))))
    return max(errs)


def w_sample(rank, world, n, seed, shots):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim

    ops_ = _mixed_circuit(n, seed)
    sv = _sharded(n, world)
    sv.apply_operations(ops_)
    got = sv.sample(shots, np.random.default_rng(seed), exact=True)
    ref = o_sim.simulate(qb.QuantumScript(ops_, [qb.sample(wires=range(n))], shots=shots),
                         rng=np.random.default_rng(seed))
    sub = sv.sample(shots, np.random.default_rng(seed + 1), wires=[2, 0], exact=True)
    # sample_state(wires=subset) (sampling.py:439-476): marginal first, then Generator.choice
    from oracle.sampling import sample_state
    st, _ = o_sim.get_final_state(qb.QuantumScript(ops_, []))
    ref_sub = sample_state(st, shots, wires=[2, 0], rng=np.random.default_rng(seed + 1))
    # device level (measure_with_samples): all wires are sampled, columns picked afterwards
    from np_engine import NumpyEngine
    from pennylane_b200.sharded import simulate_sharded
    tape = qb.QuantumScript(ops_, [qb.sample(wires=[2, 0])], shots=shots)
    dev = simulate_sharded(tape, dist, rng=np.random.default_rng(seed + 2),
                           engine=NumpyEngine(n - (world.bit_length() - 1)))
    sub_ok2 = np.array_equal(dev, o_sim.simulate(tape, rng=np.random.default_rng(seed + 2)))
    assert sub_ok2, "simulate_sharded sample(wires=subset) differs from the oracle"
    return bool(np.array_equal(got, ref)), bool(np.array_equal(sub, ref_sub)), list(sv.phys)


def w_batched(rank, world, n, seed):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q

    rng = np.random.default_rng(seed)
    B = 3
    ops_ = [q.Hadamard(wires=w) for w in range(n)]
    ops_ += [q.RY(rng.uniform(0, 6, B), wires=0), q.RZ(rng.uniform(0, 6, B), wires=0),
             q.CNOT(wires=[0, 1]), q.RX(rng.uniform(0, 6, B), wires=n - 1), q.CNOT(wires=[n - 1, 0]),
             q.RY(rng.uniform(0, 6), wires=1)]
    obs = q.PauliZ(wires=0) @ q.PauliX(wires=1)
    sv = _sharded(n, world)
    sv.apply_operations(ops_)
    got = sv.expval_pauli_sentence(obs.pauli_rep)
    ref = o_sim.simulate(qb.QuantumScript(ops_, [qb.expval(obs)]))
    return float(np.max(np.abs(np.asarray(got) - np.asarray(ref))))


def w_simulate(rank, world, n, seed):
    import pennylane_b200 as qb
    from np_engine import NumpyEngine
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q
    from pennylane_b200.sharded import simulate_sharded

    g = world.bit_length() - 1
    rng = np.random.default_rng(seed)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    ops_ = [q.StatePrep(psi, wires=range(n))] + _hea(n, 1, seed)
    tape = qb.QuantumScript(ops_, [qb.expval(q.PauliX(wires=0) @ q.PauliZ(wires=2)), qb.probs(wires=[0, 1])])
    got = simulate_sharded(tape, dist, engine=NumpyEngine(n - g))
    ref = o_sim.simulate(tape)
    return abs(got[0] - ref[0]), float(np.max(np.abs(got[1] - ref[1])))


def w_density(rank, world, n, seed):
    """Reduced density matrices and the measurements on them, global and local wires mixed."""
    import pennylane_b200 as qb
    from np_engine import NumpyEngine
    from oracle import simulate as o_sim
    from pennylane_b200.sharded import simulate_sharded

    g = world.bit_length() - 1
    ops_ = _hea(n, 2, seed)
    mps = [qb.density_matrix([0, n - 1]), qb.purity([1, 0, 3]), qb.vn_entropy([n - 2], log_base=2),
           qb.mutual_info([0], [2, 1]), qb.density_matrix([2])]
    tape = qb.QuantumScript(ops_, mps)
    got = simulate_sharded(tape, dist, engine=NumpyEngine(n - g))
    ref = o_sim.simulate(tape)
    return [float(np.max(np.abs(np.asarray(a) - np.asarray(b)))) for a, b in zip(got, ref)]


def _mcm_tape(n, seed, shots):
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.mcm import cond, measure

    m0, m1, m2 = measure(0), measure(1, reset=True), measure(n - 1, reset=True)
    ops_ = _hea(n, 2, seed) + [
        m0.measurements[0], cond(m0, q.RX(0.7, wires=0)), cond(m0, q.Hadamard(wires=n - 2)),
        q.CNOT(wires=[0, n - 1]), m1.measurements[0], cond(m0 & m1, q.RY(1.1, wires=1)),
        q.CNOT(wires=[1, 2]), m2.measurements[0], cond(~m2, q.PauliX(wires=0))]
    return qb.QuantumScript(ops_, [qb.sample(wires=range(n)), qb.sample(m0), qb.sample(m1),
                                   qb.sample(m2)], shots=shots)


def w_mcm(rank, world, n, seed, shots):
    """One-shot dynamic circuit on the sharded state vs the oracle's one-shot loop, same seed:
    measurements on a global wire (wire 0), with reset on a global and on a local wire, and
    conditionals feeding gates on global and local wires."""
    import pennylane_b200 as qb
    from np_engine import NumpyEngine
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q
    from pennylane_b200.mcm import cond, measure
    from pennylane_b200.sharded import simulate_sharded

    g = world.bit_length() - 1
    tape = _mcm_tape(n, seed, shots)
    got = simulate_sharded(tape, dist, rng=np.random.default_rng(seed), engine=NumpyEngine(n - g))
    ref = o_sim.simulate(tape, rng=np.random.default_rng(seed))
    same = all(np.array_equal(np.asarray(a[0]).reshape(-1), np.asarray(b[0]).reshape(-1))
               and [int(x) for x in a[1:]] == [int(x) for x in b[1:]] for a, b in zip(got, ref))
    return len(got), len(ref), bool(same), sorted({tuple(int(x) for x in a[1:]) for a in got})


def w_mcm_c64(rank, world, n, seed):
    """complex64 shards (round-1 advisor finding): re-measuring a wire that is deterministically
    |0> or |1> gives p = 1 + O(1e-7); the reference renormalises within 10 eps of the STATE's
    precision (apply_operation.py:451-457), so the draw must use the shard's dtype, not float64's
    eps, or it raises 'probabilities greater than 1'."""
    from np_engine import NumpyEngine
    from pennylane_b200 import ops as q
    from pennylane_b200.mcm import measure
    from pennylane_b200.sharded import ShardedStateVector

    g = world.bit_length() - 1
    sv = ShardedStateVector(n, dist, engine=NumpyEngine(n - g, dtype=np.complex64), dtype=np.complex64)
    assert sv.np_dtype == np.dtype(np.complex64)
    rng = np.random.default_rng(seed)
    outcomes = {}
    gates = _hea(n, 3, seed)
    first, again = [], []
    for w in (0, 1, n - 1):                       # wire 0 (and 1 for world 4) sit on rank bits
        m_a, m_b = measure(w), measure(w)
        gates += [m_a.measurements[0], m_b.measurements[0]]
        first.append(m_a.measurements[0]); again.append(m_b.measurements[0])
    sv.apply_gates(gates, outcomes, rng)
    return [int(outcomes[m]) for m in first], [int(outcomes[m]) for m in again], float(sv.norm2())


# ---------------------------------------------------------------------------------------------
# tests
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("kind", ["mixed", "hea"])
def test_sharded_state_matches_oracle(world, kind):
    res = run_ranks(world, "w_state", 8, kind, 11)
    for err, stats, phys in res:
        assert err < 1e-12, err
        assert stats["exchanges"] >= 1
    assert len({tuple(r[2]) for r in res}) == 1          # every rank tracks the same map


@pytest.mark.parametrize("world,piece_bits,stage_bytes", [(2, 3, 4096), (4, 2, 256), (2, 1, 1 << 20), (4, 3, 64)])
def test_overlapped_exchange_windows_match_oracle(world, piece_bits, stage_bytes):
    n = 9
    res = run_ranks(world, "w_window", n, 11, piece_bits, stage_bytes)
    ref_bytes = run_ranks(world, "w_window", n, 11, 0, stage_bytes)
    for (err, nwin, nex, partial, stats), (err0, nwin0, nex0, partial0, stats0) in zip(res, ref_bytes):
        assert err < 1e-12 and err0 < 1e-12
        assert nwin > 0 and partial > 0, "no exchange was pipelined"
        assert nwin0 == 0 and partial0 == 0, "B200Q_EXCHANGE_PIECE_BITS=0 must switch the overlap off"
        assert nwin + nex == nex0
        # cutting an exchange into pieces moves exactly the same bytes
        assert stats["exchange_bytes"] == stats0["exchange_bytes"] and stats["exchanges"] == stats0["exchanges"]


def test_piece_steps_cover_each_slab_exactly_once():
    from pennylane_b200.sharded import ShardedStateVector, _free_bit_window

    for chunk_bits, lo, pb, cap in [(10, 4, 2, 8), (10, 4, 2, 64), (10, 7, 3, 16), (12, 0, 3, 4), (8, 8, 0, 32),
                                    (8, 8, 0, 1 << 8), (9, 5, 1, 1 << 12)]:
        chunk = 1 << chunk_bits
        seen = np.zeros(chunk, dtype=int)
        for p in range(1 << pb):
            for off, run, pitch, count in ShardedStateVector._piece_steps(chunk, lo, pb, p, cap):
                assert run * count <= max(cap, run)
                for c in range(count):
                    idx = np.arange(off + c * pitch, off + c * pitch + run)
                    assert np.all(((idx >> lo) & ((1 << pb) - 1)) == p)
                    seen[idx] += 1
        assert np.all(seen == 1), (chunk_bits, lo, pb, cap)
    assert _free_bit_window(0b0000_0111, 8, 3, 0) == (5, 3)
    assert _free_bit_window(0b0110_0111, 8, 3, 0) == (3, 2)
    assert _free_bit_window(0b1010_1011, 8, 3, 2) == (6, 1)
    assert _free_bit_window(0b1111_1111, 8, 3, 0) is None


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_expval_matches_oracle(world):
    for e_h, e_z, e_n, stats in run_ranks(world, "w_expval", 8, 5):
        assert e_h < 1e-12 and e_z < 1e-12 and e_n < 1e-12, (e_h, e_z, e_n)


def test_sharded_probs_match_oracle():
    assert max(run_ranks(2, "w_probs", 7, 3)) < 1e-12
    assert max(run_ranks(4, "w_probs", 7, 3)) < 1e-12


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_samples_bit_exact(world):
    for full_ok, sub_ok, phys in run_ranks(world, "w_sample", 8, 9, 500):
        assert full_ok and sub_ok
        assert phys == list(range(8))                    # identity map restored for the CDF


def test_sharded_broadcast_parameters():
    assert max(run_ranks(2, "w_batched", 6, 2)) < 1e-12


def test_simulate_sharded_stateprep_expval_probs():
    for e0, e1 in run_ranks(2, "w_simulate", 7, 4):
        assert e0 < 1e-12 and e1 < 1e-12


def test_planner_belady_and_exchange_volume():
    """Planner only (no processes): the HEA needs one exchange per layer per rank-bit group,
    never more, and the victims are the bits used furthest in the future."""
    from pennylane_b200.sharded import ExchangeStep, RunStep, plan

    n, g = 10, 2
    steps, final = plan(_hea(n, 4, 0), n, g)
    ex = [s for s in steps if isinstance(s, ExchangeStep)]
    assert 1 <= len(ex) <= 6                             # about one per layer (4 layers)
    assert all(e.k == g for e in ex)                     # both rank bits move together
    assert sorted(final) == list(range(n))
    runs = [s for s in steps if isinstance(s, RunStep)]
    assert sum(len(r.ops) for r in runs) == len(_hea(n, 4, 0))


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_mid_circuit_measurements_match_oracle(world):
    for n_got, n_ref, same, outcomes in run_ranks(world, "w_mcm", 6, 13, 12):
        assert n_got == n_ref == 12 and same
        assert len(outcomes) > 1                       # several branches were actually visited


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_mid_measure_on_complex64_shards_renormalises(world):
    for seed in (3, 4, 5):
        for first, again, norm in run_ranks(world, "w_mcm_c64", 7, seed):
            assert first == again                     # a repeated measurement repeats its outcome
            assert abs(norm - 1.0) < 1e-5             # complex64 tolerance


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_density_measurements_match_oracle(world):
    for errs in run_ranks(world, "w_density", 7, 3):
        assert len(errs) == 5 and max(errs) < 1e-12
