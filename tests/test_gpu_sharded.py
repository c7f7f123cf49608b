"""Sharded statevector on the CUDA engine (``-m gpu``).

world size 1 runs in-process (NCCL group of one rank): it exercises ``CudaEngine`` — the C-ABI
sampler building blocks, the Pauli-term reduction, fused segments through ``compile``/``run`` —
against the oracle.  World size 2 (needs two GPUs; skipped otherwise) spawns one process per GPU
and checks the NVLink exchange path end to end: state, expectation values, bit-exact samples."""
import os
import socket
import sys
import traceback

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture(scope="module")
def dist1():
    import torch
    import torch.distributed as dist

    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0,
                                world_size=1, device_id=torch.device("cuda:0"))
        created = True
    yield dist
    if created:
        dist.destroy_process_group()


def _check_all(dist, n, seed, fusion, world_note=""):
    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200 import ops as q
    from pennylane_b200.sharded import ShardedStateVector, simulate_sharded
    from test_sharded_gloo import _hea, _mixed_circuit

    out = {}
    for kind, ops_ in (("mixed", _mixed_circuit(n, seed)), ("hea", _hea(n, 3, seed))):
        sv = ShardedStateVector(n, dist, fusion=fusion)
        sv.apply_operations(ops_)
        st, _ = o_sim.get_final_state(qb.QuantumScript(ops_, []))
        out[f"state_{kind}"] = float(np.max(np.abs(sv.to_numpy() - np.asarray(st).reshape(-1))))
        H = None
        for i in range(n - 1):
            for P in (q.PauliX, q.PauliY, q.PauliZ):
                t = P(wires=i) @ P(wires=i + 1)
                H = t if H is None else H + t
        H = H + 0.7 * (q.PauliY(wires=0) @ q.PauliZ(wires=1) @ q.PauliX(wires=n - 1))
        ref = o_sim.simulate(qb.QuantumScript(ops_, [qb.expval(H)]))
        out[f"expval_{kind}"] = abs(sv.expval_pauli_sentence(H.pauli_rep) - ref)
        out[f"norm_{kind}"] = abs(sv.norm2() - 1.0)
        for wires in ([0], [n - 1, 0, 2], [2, 3]):
            refp = o_sim.simulate(qb.QuantumScript(ops_, [qb.probs(wires=wires)]))
            out[f"probs_{kind}_{wires}"] = float(np.max(np.abs(sv.probs(wires) - refp)))
        tape = qb.QuantumScript(ops_, [qb.sample(wires=range(n))], shots=2000)
        got = simulate_sharded(tape, dist, rng=np.random.default_rng(seed), fusion=fusion)
        refs = o_sim.simulate(tape, rng=np.random.default_rng(seed))
        out[f"samples_{kind}"] = 0.0 if np.array_equal(got, refs) else 1.0
        # the state-consuming sampler (probabilities built over the state buffer in chunks)
        sv2 = ShardedStateVector(n, dist, dtype=np.complex128, fusion=fusion)
        sv2.apply_operations(ops_)
        eng = sv2.engine
        orig = eng.probs_inplace_device
        eng.probs_inplace_device = lambda: orig(chunk_bits=max(3, sv2.nl - 3))   # several chunks
        got2 = sv2.sample(2000, np.random.default_rng(seed), None, True, consume=True)
        out[f"samples_consume_{kind}"] = 0.0 if np.array_equal(got2, refs) else 1.0
    # native mid-circuit measurements (one-shot loop) on the sharded state: measured bits and
    # terminal samples of every shot equal the oracle's under the same seed
    from test_sharded_gloo import _mcm_tape

    tape = _mcm_tape(n, seed, 8)
    got = simulate_sharded(tape, dist, rng=np.random.default_rng(seed), fusion=fusion)
    ref = o_sim.simulate(tape, rng=np.random.default_rng(seed))
    same = len(got) == len(ref) == 8 and all(
        np.array_equal(np.asarray(a[0]).reshape(-1), np.asarray(b[0]).reshape(-1))
        and [int(x) for x in a[1:]] == [int(x) for x in b[1:]] for a, b in zip(got, ref))
    out["mcm_one_shot"] = 0.0 if same else 1.0
    # reduced density matrices: per-rank Gram blocks added in rank order
    ops_ = _hea(n, 2, seed)
    mps = [qb.density_matrix([0, n - 1]), qb.purity([1, 0, 3]), qb.vn_entropy([n - 2], log_base=2),
           qb.mutual_info([0], [2, 1])]
    tape = qb.QuantumScript(ops_, mps)
    got = simulate_sharded(tape, dist, fusion=fusion)
    ref = o_sim.simulate(tape)
    for i, (a, b) in enumerate(zip(got, ref)):
        out[f"density_{i}"] = float(np.max(np.abs(np.asarray(a) - np.asarray(b))))
    return out


def _check_windows(dist, n, seed):
    """Exchanges pipelined against the segments around them (sharded._schedule / _run_window):
    the state equals the oracle's and at least one window ran."""
    import pennylane_b200 as qb
    from oracle import simulate as o_sim
    from pennylane_b200.sharded import ShardedStateVector
    from test_sharded_gloo import _hea

    ops_ = _hea(n, 4, seed)
    out = {}
    for pb in (3, 1):
        os.environ["B200Q_EXCHANGE_PIECE_BITS"] = str(pb)
        try:
            sv = ShardedStateVector(n, dist, fusion=1, stage_bytes=1 << 16)
            prog = sv.compile(ops_)
            nwin = sum(1 for e in prog["schedule"] if e[0] == "window")
            sv.run(prog)
            st, _ = o_sim.get_final_state(qb.QuantumScript(ops_, []))
            out[f"state_pb{pb}"] = float(np.max(np.abs(sv.to_numpy() - np.asarray(st).reshape(-1))))
            out[f"no_window_pb{pb}"] = 0.0 if (nwin > 0 or dist.get_world_size() == 1) else 1.0
        finally:
            os.environ.pop("B200Q_EXCHANGE_PIECE_BITS", None)
    return out


@pytest.mark.parametrize("fusion", [0, 1])
def test_world1_cuda_engine_matches_oracle(dist1, fusion):
    res = _check_all(dist1, 14, 21, fusion)
    for k, v in res.items():
        assert v < 1e-12, (k, v)


def _worker(rank, world, port, q_, exchange="symm"):
    try:
        import torch
        import torch.distributed as dist

        os.environ["B200Q_EXCHANGE"] = exchange       # "nccl": the send / recv fallback of the exchange

        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world,
                                device_id=torch.device(f"cuda:{rank}"))
        res = {}
        for fusion in ((0, 1) if exchange == "symm" else (1,)):
            for k, v in _check_all(dist, 15, 33, fusion).items():
                res[f"f{fusion}_{k}"] = v
        # the overlapped schedule: specialised segment kernels forced on (their partial launches
        # are what lets a segment run piece by piece beside the exchange), 17 qubits so that
        # index bits outside the 12-bit tiles exist
        os.environ["B200Q_JIT"] = "1"
        for k, v in _check_windows(dist, 17, 5).items():
            res[f"win_{k}"] = v
        q_.put((rank, "ok", res))
        dist.destroy_process_group()
    except Exception:                                   # noqa: BLE001
        q_.put((rank, "err", traceback.format_exc()))


@pytest.mark.parametrize("world,exchange", [(2, "symm"), (4, "symm"), (2, "nccl")])
def test_multi_gpu_exchange_matches_oracle(world, exchange):
    """``exchange``: "symm" = pushes over peer-mapped symmetric memory with stream-memory-op
    flags and the TMA unpack kernel; "nccl" = the ``isend`` / ``irecv`` fallback taken when the
    symmetric allocation is refused (same window schedule, blocking pieces)."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q_ = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q_, exchange)) for r in range(world)]
    for p in procs:
        p.start()
    for _ in range(world):
        rank, status, res = q_.get(timeout=600)
        assert status == "ok", f"rank {rank}:\n{res}"
        for k, v in res.items():
            assert v < 1e-12, (rank, k, v)
    for p in procs:
        p.join(timeout=60)
