#!/usr/bin/env python
"""Mid-circuit measurement numbers (one JSON line):

* the two sweeps of one native ``MidMeasure`` at N qubits, complex128 — the single-wire
  probability reduction (``b200q_probs``, S read) and ``b200q_collapse`` (S/2 read + S written)
  — CUDA-event timed on the launching stream, against MEASURED_PEAKS.json's HBM GB/s;
* ``StateVector.reduced_dm`` (``b200q_gram_block``) for one, two and four kept wires;
* one-shot throughput (shots/s) of a dynamic circuit at M qubits — an entangling prefix, four
  measurements with conditional gates, all-wire terminal sample — with the prefix simulated once
  and collapsed branch states cached in HBM (ours) and with the prefix re-simulated every shot (what the reference's loop does,
  simulate.py:371-380), next to the oracle on the host at a smaller size.

    python tools/bench_mcm.py [N=30] [M=26] [shots=40]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pennylane_b200 import QuantumScript, measurements as M, ops  # noqa: E402
from pennylane_b200.mcm import cond, measure  # noqa: E402
from pennylane_b200.simulate import simulate  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 7700.0, "fallback (B200_PROFILING.md nominal)"


def _time(stream, fn, reps):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    fn()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev[0].record(stream)
        for _ in range(reps):
            fn()
        ev[1].record(stream)
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / reps * 1e-3


def dynamic_tape(n, shots, layers=2, seed=0):
    rng = np.random.default_rng(seed)
    gates = []
    for _ in range(layers):
        for w in range(n):
            gates += [ops.RY(rng.uniform(0, 6.28), wires=w), ops.RZ(rng.uniform(0, 6.28), wires=w)]
        gates += [ops.CNOT(wires=[w, (w + 1) % n]) for w in range(n)]
    n_prefix = len(gates)
    for k, w in enumerate([0, n // 3, n // 2, n - 1]):
        m = measure(w, reset=bool(k % 2))
        gates += [m.measurements[0], cond(m, ops.RX(0.4 + k, wires=(w + 1) % n)),
                  cond(m, ops.CNOT(wires=[(w + 1) % n, (w + 2) % n]))]
    return QuantumScript(gates, [M.sample(wires=list(range(n)))], shots=shots), n_prefix


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    Mq = int(sys.argv[2]) if len(sys.argv) > 2 else 26
    shots = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    peak, peak_src = _peak()
    out = {"what": "native mid-circuit measurement", "dtype": "c128", "peak_gbs": peak,
           "peak_source": peak_src}

    sv = StateVector(N)
    for w in range(0, N, 3):
        sv.apply_operation(ops.Hadamard(w))
    S = 16.0 * 2 ** N
    stream = torch.cuda.current_stream()
    kernels = {}
    for wire in (0, N // 2, N - 1):
        t = _time(stream, lambda: sv.probs_device([wire]), 5)
        kernels[f"probs_wire{wire}"] = {"bytes": S, "ms": t * 1e3, "gbps": S / t / 1e9,
                                        "frac": S / t / 1e9 / peak}
        # sample 0 without reset keeps the state's support: repeated launches stay comparable
        t = _time(stream, lambda: sv.collapse(wire, 0, False, 1.0), 5)
        kernels[f"collapse_wire{wire}"] = {"bytes": 1.5 * S, "ms": t * 1e3,
                                           "gbps": 1.5 * S / t / 1e9,
                                           "frac": 1.5 * S / t / 1e9 / peak}
        sv.reset()
        for w in range(0, N, 3):
            sv.apply_operation(ops.Hadamard(w))
    # reduced density matrices (b200q_gram_block): S per launch for m <= 2; m = 4 is 4 diagonal
    # blocks over S/4 and 6 off-diagonal ones over S/2 = 4 S
    for wires, nbytes in (([0], S), ([N - 1], S), ([0, N // 2], S), ([N - 2, N - 1], S),
                          ([0, 1, N // 2, N - 1], 4 * S)):
        t = _time(stream, lambda: sv.reduced_dm(wires), 3)
        kernels["reduced_dm_wires_" + "_".join(map(str, wires))] = {
            "bytes": nbytes, "ms": t * 1e3, "gbps": nbytes / t / 1e9, "frac": nbytes / t / 1e9 / peak}
    out["kernels"] = {"qubits": N, **kernels}
    del sv
    torch.cuda.empty_cache()

    tape, n_prefix = dynamic_tape(Mq, shots)
    simulate(dynamic_tape(Mq, 2)[0], rng=np.random.default_rng(0), fusion=1)     # warm-up
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    simulate(tape, rng=np.random.default_rng(1), fusion=1)
    torch.cuda.synchronize()
    t_cached = time.perf_counter() - t0
    # the reference's loop: every shot runs the whole tape
    t0 = time.perf_counter()
    for _ in range(shots):
        simulate(dynamic_tape(Mq, 1)[0], rng=np.random.default_rng(1), fusion=1)
    torch.cuda.synchronize()
    t_full = time.perf_counter() - t0
    out["one_shot"] = {"qubits": Mq, "shots": shots, "prefix_gates": n_prefix,
                       "gates_per_shot_after_first_mcm": len(tape.operations) - n_prefix,
                       "shots_per_s_prefix_once_branch_cache": shots / t_cached,
                       "shots_per_s_full_tape_per_shot": shots / t_full}

    # the same dynamic circuit through mcm_method="tree-traversal" (tree_mcm.py; reference
    # simulate.py:396-611): the outcome tree is walked once, every node splits its shot budget,
    # terminal samples are drawn once per leaf
    simulate(dynamic_tape(Mq, 2)[0], rng=np.random.default_rng(0), fusion=1, mcm_method="tree-traversal")
    torch.cuda.synchronize()
    tree = {}
    for sh in (shots, 100 * shots):
        t0 = time.perf_counter()
        simulate(dynamic_tape(Mq, sh)[0], rng=np.random.default_rng(1), fusion=1, mcm_method="tree-traversal")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tree[f"shots_{sh}"] = {"seconds": dt, "shots_per_s": sh / dt}
    out["tree_traversal"] = {"qubits": Mq, "mid_circuit_measurements": 4, "leaves_at_most": 16, **tree}

    from oracle.simulate import simulate as oracle_simulate
    n_cpu, cpu_shots = 20, 3
    t0 = time.perf_counter()
    oracle_simulate(dynamic_tape(n_cpu, cpu_shots)[0], rng=np.random.default_rng(1))
    t_cpu = time.perf_counter() - t0
    out["cpu_baseline"] = {"kind": "port", "qubits": n_cpu, "shots": cpu_shots,
                           "shots_per_s": cpu_shots / t_cpu,
                           "extrapolated_shots_per_s_at_M": cpu_shots / t_cpu / 2 ** (Mq - n_cpu),
                           "cores": os.cpu_count()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
