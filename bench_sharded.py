"""Multi-GPU arm of bench.py (`--gpus N` under torchrun): the statevector sharded over N ranks.

Weak scaling at BASELINE's multi-GPU shard size: 33 local qubits (128 GiB of complex128 per GPU),
i.e. 34 / 35 / 36 qubits on 2 / 4 / 8 GPUs (`--weak16g`: 30 + log2 N qubits, the 1-GPU shard); the
circuit is the same hardware-efficient ansatz (8 layers, expval(Z0)).  One step = reset + the whole
circuit (fused segments on every shard + the qubit-remapping exchanges over NVLink, overlapped with
the segments around them) + the expectation value.  Timed on the device with CUDA events between
barriers, max over ranks.  Besides the contract keys the line carries `parity` (a 20-qubit twin
through the same path against the oracle), `closed_form_check` (full size), the split of the step
into shard-local sweeps (HBM roofline) and exchanges (NVLink: bytes sent per GPU over the time the
communication streams were busy, against the 770 GB/s measured peer-copy figure of
B200_PROFILING.md; `visible_seconds_per_step` = what the exchanges add to the sweeps' own time).
"""
from __future__ import annotations

import json
import os
import time

import numpy as np


def _log(rank, msg, t0=[None]):
    """Progress on stderr (rank 0) — a 128 GiB-per-GPU run is long enough to want a pulse."""
    import sys
    if t0[0] is None:
        t0[0] = time.perf_counter()
    if rank == 0:
        print(f"[bench_sharded +{time.perf_counter() - t0[0]:7.1f}s] {msg}", file=sys.stderr, flush=True)


def run_sharded(args, dist, rank, world, local_rank):
    import torch

    import pennylane_b200 as qb
    from bench import ROOT, ClockSampler, hea_ops, workload_config
    from pennylane_b200 import ops as q
    from pennylane_b200.sharded import ShardedStateVector, simulate_sharded

    g = world.bit_length() - 1
    # BASELINE.json's multi-GPU point is 36 qubits on 8 GPUs: 128 GiB of state per GPU.  N = 2 and
    # 4 use the same shard size (34 / 35 qubits) so that the N > 1 points are mutually comparable;
    # `--weak16g` keeps the 1-GPU shard (30 + log2 N qubits, 16 GiB per GPU) of round 1.
    per_gpu_qubits = args.qubits if (args.weak16g or args.qubits != 30) else 33
    n = per_gpu_qubits + g
    _log(rank, f"start: {n} qubits over {world} GPUs")
    parity = parity_twin(dist, rank, world, args)
    _log(rank, f"parity twin done: {parity}")
    ops_ = hea_ops(n, args.layers)
    ngates = len(ops_)
    obs = q.PauliZ(wires=0)
    fusion = args.fusion_level if args.fusion == "on" else 0
    sv = ShardedStateVector(n, dist, dtype=np.complex128, fusion=fusion)
    _log(rank, "state allocated")
    program = sv.compile(ops_)
    _log(rank, f"compiled: {program['n_exchanges']} exchanges")
    S_loc = 16.0 * (1 << sv.nl)

    records = []

    def timer(kind, fn):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        records.append((kind, e0, e1, r if kind in ("run", "window") else 0))
        return r

    def step():
        sv.reset()
        sv.run(program)
        return sv.expval_pauli_sentence(obs.pauli_rep)

    for i in range(args.warmup):
        step()
        torch.cuda.synchronize()
        _log(rank, f"warm-up step {i} done")
    torch.cuda.synchronize()
    dist.barrier()
    clocks = ClockSampler(local_rank)
    clocks.start()
    sv.timer = timer
    sv.stats = {k: 0 for k in sv.stats}
    start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    start.record()
    val = None
    for _ in range(args.steps):
        val = step()
    end.record()
    torch.cuda.synchronize()
    dist.barrier()
    sv.timer = None
    _log(rank, "timed steps done")
    ms = torch.tensor([start.elapsed_time(end)], dtype=torch.float64, device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    ms_per_step = total_ms / args.steps
    clk = clocks.stop()

    # "run": sweeps outside any window; "window": [segments] + exchange + [segments] pipelined piece
    # by piece (sharded._schedule); "exchange": an exchange nothing could be overlapped with
    run_s = sum(a.elapsed_time(b) for k, a, b, _ in records if k == "run") * 1e-3
    ex_s = sum(a.elapsed_time(b) for k, a, b, _ in records if k == "exchange") * 1e-3
    win_s = sum(a.elapsed_time(b) for k, a, b, _ in records if k == "window") * 1e-3
    run_sweeps = sum(c for k, _, _, c in records if k == "run")
    win_sweeps = sum(c for k, _, _, c in records if k == "window")
    sweeps = run_sweeps + win_sweeps
    ex_bytes = sv.stats["exchange_bytes"]
    # time the communication stream spent on the pieces (pack + barrier + pull over NVLink)
    comm_s = sum(a.elapsed_time(b) for a, b in getattr(sv, "comm_records", [])) * 1e-3
    sweep_s = run_s / run_sweeps if run_sweeps else 0.0
    # the part of the exchanges that is NOT hidden behind sweeps: whole un-overlapped exchanges
    # plus what the windows take beyond their own sweeps at the pace of the sweeps outside
    ex_visible_s = ex_s + max(0.0, win_s - win_sweeps * sweep_s)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm = run_sweeps * 2.0 * S_loc / run_s / 1e9 if run_s > 0 else None
    n_windows = sum(1 for k, *_ in records if k == "window") // args.steps
    n_plain = sum(1 for k, *_ in records if k == "exchange") // args.steps

    # Closed-form check at FULL size through the same sharded path (planner, overlapped exchanges,
    # partial segment launches): RY(theta_w) on every wire, then a CNOT chain 0 -> 1 -> ... -> n-1,
    # <Z_k> = prod_{j <= k} cos(theta_j); wire 0 .. g-1 start on the rank bits, so the chain forces
    # exchanges.  Nothing else pins correctness above the ~24 qubits the oracle can hold.
    closed = None
    if fusion and not args.quick:
        th = np.random.default_rng(11).uniform(0.2, 1.2, n)
        cops = [q.RY(float(th[w]), wires=w) for w in range(n)] + [q.CNOT(wires=[w, w + 1]) for w in range(n - 1)]
        sv.reset()
        ex0 = sv.stats["exchanges"]
        sv.run(sv.compile(cops))
        errs = {}
        for k in sorted({0, 1, n // 2, n - 2, n - 1}):
            got = float(sv.expval_pauli_sentence(q.PauliZ(wires=k).pauli_rep))
            errs[k] = abs(got - float(np.prod(np.cos(th[: k + 1]))))
        closed = {"circuit": f"RY(theta_w) on {n} wires + CNOT chain, <Z_k> = prod_(j<=k) cos(theta_j)",
                  "wires_checked": sorted(errs), "max_abs_err": max(errs.values()),
                  "norm2_minus_1": float(sv.norm2()) - 1.0, "exchanges": sv.stats["exchanges"] - ex0}
        assert closed["max_abs_err"] < 1e-12, closed
        _log(rank, f"closed-form check done: {closed}")

    # e2e: the public sharded entry point, host parameters in / host scalar out, wall clock
    par = np.random.default_rng(3).uniform(0, 2 * np.pi, (args.layers, n, 2))

    def e2e_step():
        ops2 = []
        for l in range(args.layers):
            for w in range(n):
                ops2.append(q.RY(float(par[l, w, 0]), wires=w))
                ops2.append(q.RZ(float(par[l, w, 1]), wires=w))
            for w in range(n):
                ops2.append(q.CNOT(wires=[w, (w + 1) % n]))
        tape = qb.QuantumScript(ops2, [qb.expval(q.PauliZ(wires=0))])
        return simulate_sharded(tape, dist, fusion=fusion)

    sv_symm = getattr(sv, "_symm", None)
    del sv, program
    torch.cuda.empty_cache()
    e2e_step()
    torch.cuda.synchronize(); dist.barrier()
    _log(rank, "first e2e step done")
    k2 = max(1, min(args.steps, 2))
    t0 = time.perf_counter()
    for _ in range(k2):
        r = e2e_step()
    torch.cuda.synchronize(); dist.barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / k2], dtype=torch.float64, device="cuda")
    dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s.item())
    assert abs(float(r) - float(val)) < 1e-9, (r, val)

    if rank == 0:
        cfg = workload_config(args, n)
        cfg["shard_bytes"] = int(S_loc)
        cfg["workload"] += (f"; sharded over {world} GPUs by the top {g} qubits, {per_gpu_qubits} local qubits "
                            f"({int(S_loc) >> 30} GiB) per GPU" +
                            ("" if per_gpu_qubits == 33 else " [--weak16g / --qubits: not BASELINE's 36q@8 shard size]"))
        line = {
            "metric": "gates_per_s", "value": ngates / (ms_per_step * 1e-3), "unit": "gates/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
            "data": "synthetic", "config": cfg, "expval": float(val),
            # gates/s falls by 2 per added qubit at fixed hardware speed; the size-independent
            # figure is amplitude updates per second and per GPU
            "amplitude_updates_per_s_per_gpu": ngates * float(2 ** n) / world / (ms_per_step * 1e-3),
            "parity": parity, "closed_form_check": closed,
            "state_sweeps_per_step": sweeps // args.steps,
            "roofline": {"bound": "hbm", "kernel": "sk_kernel (csrc/segk.cuh, structure-specialised fused segment on each shard: 2*S_loc per launch)",
                         "achieved": hbm, "peak": peak, "unit": "GB/s",
                         "frac": (hbm / peak) if hbm else None, "traffic": None,
                         "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                         "launches_timed_alone": run_sweeps // args.steps,
                         "share_of_step": (run_s + win_sweeps * sweep_s) / (total_ms * 1e-3),
                         "all_sweeps_gbps_incl_exchange_stalls": sweeps * 2.0 * S_loc / (total_ms * 1e-3) / 1e9},
            "exchange": {"per_step": sv_stats_per_step(ex_bytes, args.steps),
                         "count_per_step": n_windows + n_plain,
                         "overlapped_with_sweeps": n_windows, "not_overlapped": n_plain,
                         "mode": "copy-engine pushes into the partners' staging buffers (peer-mapped symmetric memory), "
                                 "stream-memory-op flags, TMA unpack kernel; two communication streams, pipelined "
                                 "piece by piece against the segments before / after (sharded._schedule)" if sv_symm
                                 else "NCCL send/recv (no overlap)",
                         "comm_stream_seconds_per_step": comm_s / args.steps if comm_s else None,
                         "sent_gbps_per_gpu": (ex_bytes / comm_s / 1e9) if comm_s > 0 else
                                              (ex_bytes / ex_s / 1e9 if ex_s > 0 else None),
                         "nvlink_peak_gbps": 770.0, "peak_source": "B200_PROFILING.md measured peer copy, per direction",
                         "frac": (ex_bytes / comm_s / 1e9 / 770.0) if comm_s > 0 else
                                 ((ex_bytes / ex_s / 1e9 / 770.0) if ex_s > 0 else None),
                         "visible_seconds_per_step": ex_visible_s / args.steps,
                         "share_of_step": ex_visible_s / (total_ms * 1e-3)},
            "cpu_baseline": None,
            "e2e": {"value": ngates / e2e_s, "unit": "gates/s", "seconds_per_step": e2e_s,
                    "h2d_bytes_per_step": int(world * 64 * ngates), "d2h_bytes_per_step": 8 * world},
            "gpu_launches": int((sweeps + 3 * args.steps) * world), "clocks": clk,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def parity_twin(dist, rank, world, args, n=20, shots=2000):
    """A 20-qubit twin of the workload through the SAME sharded path (planner, exchanges over
    NCCL, specialised segment kernels forced on) against the oracle, which rank 0 runs on the
    host: final state, expectation value and seeded samples.  Reported in the bench line."""
    import torch

    import pennylane_b200 as qb
    from bench import hea_ops
    from pennylane_b200 import ops as q
    from pennylane_b200.sharded import ShardedStateVector

    prev = os.environ.get("B200Q_JIT")
    os.environ["B200Q_JIT"] = "1"
    try:
        ops_ = hea_ops(n, 3, seed=5)
        fusion = args.fusion_level if args.fusion == "on" else 0
        sv = ShardedStateVector(n, dist, dtype=np.complex128, fusion=fusion)
        sv.run(sv.compile(ops_))
        ev = float(sv.expval_pauli_sentence((q.PauliZ(wires=0) @ q.PauliX(wires=n - 1)).pauli_rep))
        n_ex = sv.stats["exchanges"]
        state = sv.to_numpy()
        samples = sv.sample(shots, np.random.default_rng(17), exact=True)
        out = None
        if rank == 0:
            from oracle import simulate as o_sim
            tape = qb.QuantumScript(ops_, [qb.state()])
            ref, _ = o_sim.get_final_state(tape)
            ref = np.asarray(ref).reshape(-1)
            tape_e = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0) @ q.PauliX(wires=n - 1))])
            ref_e = float(o_sim.measure_final_state(tape_e, o_sim.get_final_state(tape_e)[0], False))
            tape_s = qb.QuantumScript(ops_, [qb.sample(wires=range(n))], shots=shots)
            ref_s = o_sim.simulate(tape_s, rng=np.random.default_rng(17))
            out = {"circuit": f"hea{n} (3 layers, seed 5) sharded over {world} GPUs, {n_ex} exchanges",
                   "max_abs_err_state": float(np.max(np.abs(np.asarray(state).reshape(-1) - ref))),
                   "abs_err_expval": abs(ev - ref_e),
                   "samples_identical": bool(np.array_equal(np.asarray(samples), np.asarray(ref_s))),
                   "shots": shots, "oracle": "oracle/ (numpy restatement of default.qubit) on rank 0"}
            assert out["max_abs_err_state"] < 1e-12 and out["abs_err_expval"] < 1e-12, out
        del sv
        torch.cuda.empty_cache()
        return out
    finally:
        if prev is None:
            os.environ.pop("B200Q_JIT", None)
        else:
            os.environ["B200Q_JIT"] = prev


def sv_stats_per_step(nbytes, steps):
    return {"bytes_sent_per_gpu": int(nbytes // max(1, steps))}
