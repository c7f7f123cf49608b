#!/usr/bin/env python
"""Run BASELINE.json's configurations at (or near) full size on one B200 through the public
device API and print one JSON line per configuration (wall-clock seconds, gates/s, checks).

  python tools/run_configs.py [c1] [c2] [c4] [c5] [--c4-qubits 31] [--c5-qubits 30]

c1 is checked against the oracle at full size (20 qubits); c2/c4/c5 are timed at full size and
checked through size-independent properties (norm, variational bounds, sample statistics): their
oracle parity lives in tests/ on shrunk twins.  Not part of bench.py.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_b200 as qb  # noqa: E402
from pennylane_b200 import ops as q  # noqa: E402


def sync():
    import torch
    torch.cuda.synchronize()


LAST = {}


def timed(fn, reps=1, warm=2):
    from pennylane_b200.statevector import SWEEPS
    for _ in range(warm):     # warm-up: the 2nd call of a structure promotes it to the compiled path
        fn()
    sync()
    before = dict(SWEEPS)
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    sync()
    dt = (time.perf_counter() - t0) / reps
    LAST["sweeps"] = {k: (SWEEPS[k] - before[k]) / reps for k in SWEEPS}
    return out, dt


def roofline(state_bytes, seconds, extra_reads=0.0):
    """Whole-call figure: (fused segment launches + per-gate kernels) x 2 S + read-only sweeps,
    over the WALL time of the call (host work included), against the measured HBM peak."""
    peak = 6547.5
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    sw = LAST.get("sweeps", {})
    n = sw.get("fused_segments", 0) + sw.get("per_gate", 0)
    gb = (2.0 * n + extra_reads) * state_bytes / seconds / 1e9
    return {"state_sweeps": sw, "bytes_credited": "2*S per sweep" + (f" + {extra_reads:g}*S of reductions" if extra_reads else ""),
            "gbps_over_wall_time": gb, "frac_of_hbm_peak": gb / peak, "peak": peak}


def c1():
    from oracle import adjoint_jacobian as o_adj
    from oracle import simulate as o_sim
    n, layers = 20, 4
    w = np.random.default_rng(1).uniform(0, 2 * np.pi, (layers, n, 3))
    ops_ = []
    for l in range(layers):
        for i in range(n):
            ops_.append(q.Rot(w[l, i, 0], w[l, i, 1], w[l, i, 2], wires=i))
        r = (l % (n - 1)) + 1
        for i in range(n):
            ops_.append(q.CNOT(wires=[i, (i + r) % n]))
    tape = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0))])
    dev = qb.B200Qubit(wires=n)
    (ptape,), config = dev.preprocess(tape, qb.ExecutionConfig(gradient_method="adjoint"))
    (res, jac), dt = timed(lambda: dev.execute_and_compute_derivatives(ptape, config), 3)
    t0 = time.perf_counter()
    st, _ = o_sim.get_final_state(ptape)
    ref = o_sim.measure_final_state(ptape, st, False)
    ref_jac = np.array(o_adj.adjoint_jacobian(ptape, st), dtype=float)
    t_cpu = time.perf_counter() - t0
    return {"config": "c1: 20q StronglyEntanglingLayers x4, expval(Z0) + adjoint (240 params)",
            "roofline": dict(roofline(16.0 * (1 << n), dt), note="16 MiB state: launch / host bound, not HBM bound"),
            "seconds": dt, "oracle_seconds": t_cpu, "speedup_vs_oracle": t_cpu / dt,
            "expval_abs_err": abs(float(res) - float(ref)),
            "jacobian_max_abs_err": float(np.max(np.abs(np.array(jac, dtype=float) - ref_jac)))}


def c2():
    import networkx as nx
    n, p = 26, 4
    g = nx.random_regular_graph(3, n, seed=2)
    edges = list(g.edges)
    par = np.random.default_rng(2).uniform(0, 2 * np.pi, (2, p))
    ops_ = [q.Hadamard(wires=i) for i in range(n)]
    for l in range(p):
        for a, b in edges:
            ops_.append(q.PauliRot(par[0, l], "ZZ", wires=[a, b]))
        for i in range(n):
            ops_.append(q.PauliRot(2 * par[1, l], "X", wires=[i]))
    cost = q.LinearCombination(
        [0.5] * len(edges) + [-0.5] * len(edges),
        [q.PauliZ(wires=a) @ q.PauliZ(wires=b) for a, b in edges] + [q.Identity(wires=a) for a, _ in edges])
    tape = qb.QuantumScript(ops_, [qb.expval(cost)])
    out = {"config": f"c2: 26q QAOA MaxCut p=4, {len(edges)} edges, {len(ops_)} gates, expval(cost_h)"}
    vals = {}
    for name, dtype in (("c128", np.complex128), ("c64", np.complex64)):
        dev = qb.B200Qubit(wires=n, c_dtype=dtype, fusion=1)
        val, dt = timed(lambda: dev.execute(tape), 3)
        vals[name] = float(val)
        out[name] = {"seconds": dt, "gates_per_s": len(ops_) / dt, "expval": float(val),
                     "roofline": roofline(float(np.dtype(dtype).itemsize) * (1 << n), dt, extra_reads=1.0)}
    out["c64_vs_c128_rel"] = abs(vals["c64"] - vals["c128"]) / max(1.0, abs(vals["c128"]))
    out["bounds_ok"] = bool(-len(edges) <= vals["c128"] <= 0.0)      # cost_h spectrum is [-|E|, 0]
    return out


def _heisenberg(n):
    coeffs, obs = [], []
    for i in range(n - 1):
        for P in (q.PauliX, q.PauliY, q.PauliZ):
            coeffs.append(1.0)
            obs.append(P(wires=i) @ P(wires=i + 1))
    return q.LinearCombination(coeffs, obs)


def c4(n, batch, dtype):
    layers = 8
    par = np.random.default_rng(4).uniform(0, 2 * np.pi, (layers, n, 2, batch))
    ops_ = []
    for l in range(layers):
        for w in range(n):
            ops_.append(q.RY(par[l, w, 0], wires=w))
            ops_.append(q.RZ(par[l, w, 1], wires=w))
        for w in range(n):
            ops_.append(q.CNOT(wires=[w, (w + 1) % n]))
    H = _heisenberg(n)
    tape = qb.QuantumScript(ops_, [qb.expval(H)])
    dev = qb.B200Qubit(wires=n, c_dtype=dtype, fusion=1)
    val, dt = timed(lambda: dev.execute(tape), 1)       # second call: compiled structures cached
    val = np.asarray(val, dtype=float)
    return {"config": f"c4: {n}q HEA x8 with parameter broadcast B={batch} ({np.dtype(dtype).name}), "
                      f"expval(Heisenberg chain, {3 * (n - 1)} Pauli words)",
            "seconds": dt, "gates_per_s": batch * len(ops_) / dt, "expval": val.tolist(),
            "roofline": roofline(float(batch * np.dtype(dtype).itemsize) * (1 << n), dt),
            "bounds_ok": bool(np.all(np.abs(val) <= 3 * (n - 1) + 1e-9)),
            "state_bytes": int(batch * np.dtype(dtype).itemsize * (1 << n))}


def c5(n, shots):
    depth = 20
    rng = np.random.default_rng(5)
    ops_ = []
    for _ in range(depth):
        for i in range(n):
            ops_.append(q.Rot(*rng.uniform(0, 2 * np.pi, 3), wires=i))
        perm = rng.permutation(n)
        for a, b in zip(perm[::2], perm[1::2]):
            ops_.append(q.CNOT(wires=[int(a), int(b)]))
    tape = qb.QuantumScript(ops_, [qb.sample(wires=range(n))], shots=shots)
    s2 = np.asarray(qb.B200Qubit(wires=n, seed=5, fusion=1).execute(tape))     # warm-up (compiles)
    sync()
    dev = qb.B200Qubit(wires=n, seed=5, fusion=1)
    from pennylane_b200.statevector import SWEEPS
    before = dict(SWEEPS)
    t0 = time.perf_counter()
    s = dev.execute(tape)
    sync()
    dt = time.perf_counter() - t0
    LAST["sweeps"] = {k: SWEEPS[k] - before[k] for k in SWEEPS}
    s = np.asarray(s)
    # the circuit alone (state only), to split the time into sweeps and sampling
    dev_c = qb.B200Qubit(wires=n, seed=5, fusion=1)
    tape_c = qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0))])
    _, dt_c = timed(lambda: dev_c.execute(tape_c), 1)
    return {"config": f"c5 (1 GPU): {n}q random circuit depth 20 ({len(ops_)} gates) + {shots} shots, seed 5",
            "seconds": dt, "circuit_seconds": dt_c, "sampling_seconds": dt - dt_c,
            "roofline_circuit": roofline(16.0 * (1 << n), dt_c, extra_reads=1.0),
            "shape": list(s.shape), "ones_fraction": float(s.mean()),
            "same_seed_identical": bool(np.array_equal(s, s2)),
            "distinct_bitstrings": int(len(np.unique(s @ (1 << np.arange(n)[::-1].astype(np.int64))))) if n <= 62 else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("which", nargs="*", default=["c1", "c2", "c4", "c5"])
    ap.add_argument("--c4-qubits", type=int, default=32)
    ap.add_argument("--c4-batch", type=int, default=2)
    ap.add_argument("--c4-dtype", default="c128")
    ap.add_argument("--c5-qubits", type=int, default=30)
    ap.add_argument("--c5-shots", type=int, default=1000000)
    a = ap.parse_args()
    for w in a.which:
        try:
            if w == "c1": r = c1()
            elif w == "c2": r = c2()
            elif w == "c4": r = c4(a.c4_qubits, a.c4_batch, np.complex128 if a.c4_dtype == "c128" else np.complex64)
            elif w == "c5": r = c5(a.c5_qubits, a.c5_shots)
            else: continue
        except Exception as exc:           # report and carry on with the next configuration
            r = {"config": w, "error": f"{type(exc).__name__}: {exc}"}
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
