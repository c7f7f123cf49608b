#!/usr/bin/env python
"""Benchmark of the hot path on B200 — one JSON line per run (contract in the task statement).

Workload (BASELINE.json `metric`: "gates/s + HBM GB/s at 30q (1 GPU) ...; adjoint Jacobian
s/step"; `configs[2]`): the 30-qubit hardware-efficient ansatz in complex128,
8 layers x (RY, RZ on every wire + CNOT ring) = 480 parameters / 720 gates, expval(Z0),
parameters `default_rng(3).uniform(0, 2pi)`.

  step      = one forward execution of the 720-gate circuit + the expectation value.
  value     = gates/s with everything resident on the device (CUDA events around the launches).
  e2e       = gates/s through `B200Qubit.execute(tape)`: host parameters in, np.float64 out,
              wall clock between device synchronisations (gate matrices travel host->device
              every step, the result comes back every step).
  roofline  = the dominant kernel family by device time in the timed region (per-gate CUDA
              events), algorithmic bytes 2*f*S per launch.
  adjoint   = seconds per execute_and_compute_derivatives step of the same tape (extra key).
  cpu_baseline = the oracle (numpy restatement of default.qubit) on the host cores, on a bounded
              sample: a 26-qubit twin of the circuit, extrapolated x2 per qubit (stated).

`--impl reference` times that CPU restatement alone (the reference itself cannot be installed:
autograd / autoray / rustworkx are absent from the image).
`--gpus N` (N > 1, under torchrun): the statevector is sharded over N ranks by its top qubits;
weak scaling (30 + log2 N qubits, 16 GiB per GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# --------------------------------------------------------------------------------------------
# workload
# --------------------------------------------------------------------------------------------
def hea_ops(n, layers=8, seed=3):
    """Hardware-efficient ansatz: per layer RY, RZ on every wire then a CNOT ring."""
    from pennylane_b200 import ops as q

    par = np.random.default_rng(seed).uniform(0, 2 * np.pi, (layers, n, 2))
    ops_ = []
    for l in range(layers):
        for w in range(n):
            ops_.append(q.RY(par[l, w, 0], wires=w))
            ops_.append(q.RZ(par[l, w, 1], wires=w))
        for w in range(n):
            ops_.append(q.CNOT(wires=[w, (w + 1) % n]))
    return ops_


def hea_tape(n, layers=8, seed=3):
    import pennylane_b200 as qb
    from pennylane_b200 import ops as q

    ops_ = hea_ops(n, layers, seed)
    return qb.QuantumScript(ops_, [qb.expval(q.PauliZ(wires=0))],
                            trainable_params=list(range(2 * n * layers)))


def credited_bytes(op, n, itemsize=16):
    """Algorithmic bytes of one gate sweep: 2 * f * S."""
    S = itemsize * (1 << n)
    f = 1.0
    if op.name in ("CNOT", "CZ", "CY", "CRX", "CRY", "CRZ", "PauliZ", "S", "T", "PhaseShift"):
        f = 0.5
    elif op.name in ("Toffoli", "CCZ", "ControlledPhaseShift"):
        f = 0.25
    return 2.0 * f * S


# --------------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=3)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
            except ValueError:
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(pw)) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------
# CPU baseline (oracle) — bounded sample
# --------------------------------------------------------------------------------------------
def cpu_baseline_sample(n_full, n_twin=26, gates_per_kind=8, repeat=1):
    """Time the oracle's apply_operation on a twin of the workload: RY/RZ on `gates_per_kind`
    evenly spread wires and as many ring CNOTs, complex128, best of `repeat`.  Returns
    (gates_per_s at n_full [extrapolated x2 per qubit], description, cores)."""
    from oracle.apply_operation import apply_operation
    from pennylane_b200 import ops as q

    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16 << 30
    while n_twin > 20 and 6 * 16 * (1 << n_twin) > avail:
        n_twin -= 1
    cores = os.cpu_count() or 1
    # torchrun exports OMP_NUM_THREADS=1: give numpy's BLAS / OpenMP pools every host core back
    # for the CPU arm (the reference's numpy kernels use whatever the pools offer)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    rng = np.random.default_rng(3)
    wires = [int(round(x)) for x in np.linspace(0, n_twin - 1, gates_per_kind)]
    ops_ = []
    for w in wires:
        ops_ += [q.RY(rng.uniform(0, 6), wires=w), q.RZ(rng.uniform(0, 6), wires=w)]
    for w in wires:
        ops_.append(q.CNOT(wires=[w, (w + 1) % n_twin]))
    state = np.zeros((2,) * n_twin, dtype=np.complex128)
    state[(0,) * n_twin] = 1.0
    state = apply_operation(q.Hadamard(wires=0), state)       # warm-up, touches all pages
    best = None
    for _ in range(repeat):
        st = state
        t0 = time.perf_counter()
        for op in ops_:
            st = apply_operation(op, st)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    gps_twin = len(ops_) / best
    scale = 2.0 ** (n_full - n_twin)
    desc = (f"oracle (numpy restatement of default.qubit apply_operation) on a {n_twin}-qubit twin: "
            f"{len(ops_)} gates (RY, RZ, ring CNOT on {gates_per_kind} spread wires) in {best:.2f} s "
            f"= {gps_twin:.3f} gates/s; extrapolated x2 per qubit (/{scale:.0f}) to {n_full} qubits")
    return gps_twin / scale, desc, cores, best, n_twin


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_gpu = args.qubits if (args.gpus == 1 or args.weak16g or args.qubits != 30) else 33
    n_full = per_gpu + int(np.log2(args.gpus))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        v, desc, cores, dt, n_twin = cpu_baseline_sample(n_full, gates_per_kind=4 if args.quick else 8)
        if i >= args.warmup:
            vals.append((v, dt))
        last = (desc, cores, n_twin)
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": "gates_per_s", "value": value, "unit": "gates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic",
        "config": dict(workload_config(args, n_full),
                       measured_on=f"{last[2]}-qubit twin of the circuit (RY, RZ, ring CNOT on spread wires), "
                                   "gates/s divided by 2 per missing qubit (BASELINE.md section 3); "
                                   "ms_per_step is the twin's time",
                       reference_twin_qubits=last[2], extrapolation_factor=float(2.0 ** (n_full - last[2]))),
        "cpu_baseline": {"value": value, "unit": "gates/s", "cores": last[1], "kind": "port",
                         "sample": last[0]},
        "e2e": {"value": value, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def ncu_traffic(family, n):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the
    committed `ncu --set full` capture of that kernel at this size (profiles/ncu_traffic.json,
    written from the .ncu-rep by tools/ncu_summary.py); None when no capture matches."""
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return float(table[family][str(n)]["dram_bytes_per_launch"])
    except Exception:
        return None


def workload_config(args, n):
    return {"workload": f"hea{n}_c128: {n}-qubit hardware-efficient ansatz, {args.layers} layers x "
                        f"(RY,RZ per wire + CNOT ring) = {2 * n * args.layers} params / "
                        f"{3 * n * args.layers} gates, expval(Z0), complex128",
            "qubits": n, "layers": args.layers, "gates": 3 * n * args.layers,
            "state_bytes": 16 * (1 << n), "l2_policy": "inputs larger than L2 (state >> 126 MB)"
            if n >= 24 else "state fits L2: flushed between steps",
            "fusion": f"{args.fusion} (level {args.fusion_level})" if args.fusion == "on" else "off", "parallelism": f"shard{args.gpus}" if args.gpus > 1 else "single"}


# --------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.simulate import measure
    from pennylane_b200.statevector import StateVector

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    if world > 1:
        from bench_sharded import run_sharded   # noqa: WPS433 (multi-GPU path lives beside us)

        return run_sharded(args, dist, rank, world, local_rank)

    n = args.qubits
    tape = hea_tape(n, args.layers)
    ops_ = tape.operations
    ngates = len(ops_)
    mp = tape.measurements[0]
    sv = StateVector(n, dtype=np.complex128)
    flush = None
    if 16 * (1 << n) < (256 << 20):
        flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    fam = {}          # kernel family -> [events...]
    launches = 0

    fused_segments = None
    use_program = False
    if args.fusion == "on":
        from pennylane_b200 import program as _program
        prog, _ = _program.get_program(sv, ops_, args.fusion_level)
        if prog is not None:
            fused_segments, use_program = prog.segs, True      # specialised kernels (segk.cuh)
        else:
            fused_segments = sv.compile_fused(ops_, level=args.fusion_level)

    def forward(record):
        nonlocal launches
        sv.reset()
        launches += 2
        if fused_segments is not None:
            S2 = 2.0 * 16 * (1 << n)
            if use_program:
                # what execute() does on a structure it has seen: rebind the parameter values
                # (host, ~4 ms, overlapped with the device) and launch the cached kernels
                pg, _hit = _program.get_program(sv, ops_, args.fusion_level)
                items = list(zip(pg.segs, pg.plans, pg.tables))
            else:
                items = [(seg, None, None) for seg in fused_segments]
            for seg, plan, tab in items:
                if record:
                    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                    e0.record()
                if plan is not None:
                    sv._launch_plan(plan, tab)
                else:
                    sv.run_segment(seg)
                launches += 1
                if record:
                    e1.record()
                    fam.setdefault("tile_segment", []).append((e0, e1, S2))
        for op in (ops_ if fused_segments is None else ()):
            if record:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            sv.apply_operation(op)
            launches += 1
            if record:
                e1.record()
                fam.setdefault(op.name, []).append((e0, e1, credited_bytes(op, n)))
        if record:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
        val = measure(mp, sv)
        launches += 2
        if record:
            e1.record()
            fam.setdefault("expval", []).append((e0, e1, 16.0 * (1 << n)))
        return val

    for _ in range(args.warmup):
        if flush is not None:
            flush.zero_()
        forward(False)
    torch.cuda.synchronize()
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches = 0
    start = torch.cuda.Event(enable_timing=True); end = torch.cuda.Event(enable_timing=True)
    step_ms = []
    val = None
    torch.cuda.synchronize()
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()
        start.record()
        val = forward(True)
        end.record()
        torch.cuda.synchronize()
        step_ms.append(start.elapsed_time(end))
    total_ms = float(np.sum(step_ms))
    ms_per_step = total_ms / args.steps
    value = ngates / (ms_per_step * 1e-3)
    timed_launches = launches

    # roofline: dominant family by device time
    fam_stats = {}
    for name, evs in fam.items():
        t = sum(a.elapsed_time(b) for a, b, _ in evs) * 1e-3
        byt = sum(c for _, _, c in evs)
        fam_stats[name] = {"launches": len(evs), "seconds": t, "bytes": byt,
                           "gbps": byt / t / 1e9 if t > 0 else None}
    dom = max((k for k in fam_stats if k != "expval"), key=lambda k: fam_stats[k]["seconds"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    kernel_of = {"tile_segment": ("sk_kernel (csrc/segk.cuh: structure-specialised fused segment, NVRTC; "
                                  "2*S per launch)" if use_program else
                                  "k_rtile<double,4,1,256,2> (fused segment: 2*S per launch)"),
                 "RY": "k_dense<double,1,byval>", "RZ": "k_parity_phase<double>",
                 "CNOT": "k_dense<double,1,byval> (1 control)"}
    all_gate_bytes = sum(v["bytes"] for k, v in fam_stats.items() if k != "expval")
    all_gate_secs = sum(v["seconds"] for k, v in fam_stats.items() if k != "expval")
    roofline = {"bound": "hbm", "kernel": kernel_of.get(dom, dom),
                "achieved": fam_stats[dom]["gbps"], "peak": peak, "unit": "GB/s",
                "frac": fam_stats[dom]["gbps"] / peak, "traffic": ncu_traffic(dom, n), "peak_source": peak_src,
                "share_of_step": fam_stats[dom]["seconds"] / (total_ms * 1e-3),
                "per_family": {k: {"launches": v["launches"], "gbps": v["gbps"],
                                   "frac": (v["gbps"] or 0) / peak,
                                   "ms_per_launch": 1e3 * v["seconds"] / v["launches"]}
                               for k, v in fam_stats.items()},
                "all_gates_gbps": all_gate_bytes / all_gate_secs / 1e9}

    # The per-gate kernels (the reference's "one sweep per gate", simulate.py:214-235) beside the
    # fused one: a few unfused launches of each gate family on wires spread over the register,
    # timed outside the step (they are the path's fusion=0 parity mode and the sharded fallback).
    if fused_segments is not None and not args.quick:
        per_gate = {}
        wires = sorted({0, n // 3, n // 2, (2 * n) // 3, n - 1})
        probes = {"RY": [q.RY(0.3 + w, wires=w) for w in wires],
                  "RZ": [q.RZ(0.7 + w, wires=w) for w in wires],
                  "CNOT": [q.CNOT(wires=[w, (w + 1) % n]) for w in wires]}
        for name, plist in probes.items():
            for op in plist:
                sv.apply_operation(op)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for op in plist:
                sv.apply_operation(op)
            e1.record()
            torch.cuda.synchronize()
            byt = sum(credited_bytes(op, n) for op in plist)
            gb = byt / (e0.elapsed_time(e1) * 1e-3) / 1e9
            per_gate[name] = {"kernel": kernel_of[name], "launches": len(plist), "gbps": gb, "frac": gb / peak}
            # per wire position (the bit-0 cases): a CNOT whose CONTROL is index bit 0
            # touches every other 16-byte amplitude, DRAM still moves whole 32-byte sectors, so
            # that position can reach at most half of its credited roofline
            per_wire = {}
            for op in plist:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                sv.apply_operation(op); sv.apply_operation(op)
                e1.record()
                torch.cuda.synchronize()
                per_wire[str(list(op.wires))] = round(
                    2 * credited_bytes(op, n) / (e0.elapsed_time(e1) * 1e-3) / 1e9 / peak, 3)
            per_gate[name]["frac_per_wires"] = per_wire
            if name == "CNOT":
                # control on index bit 0: the touched amplitudes are every other 16 bytes, so every
                # 32-byte DRAM sector is read and written: the bytes that move are 2*S, not S
                key = str(list(plist[-1].wires))
                per_gate[name]["bit0_control_frac_at_sector_level"] = {key: round(2 * per_wire[key], 3)}
        # Wide dense blocks (apply_operation.py:202-255, the tensordot path): compute-bound from
        # K = 5 on — 4 * 2^K DFMA per amplitude against 32 bytes — so their roofline is the FP64
        # pipe (34.1 TFLOP/s measured with tools/micro/fp64_forms.cu, profiles/r2_fp64_forms.txt).
        fp64_peak = 34.1e12
        rng_u = np.random.default_rng(5)
        for K in (4, 6, 8):
            m = np.linalg.qr(rng_u.normal(size=(1 << K, 1 << K)) + 1j * rng_u.normal(size=(1 << K, 1 << K)))[0]
            wires_k = sorted({int(round(x)) for x in np.linspace(1, n - 2, K)})
            while len(wires_k) < K:
                wires_k = sorted(set(wires_k) | {len(wires_k)})
            op = q.QubitUnitary(m, wires=wires_k[:K])
            sv.apply_operation(op)
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            sv.apply_operation(op)
            e1.record()
            torch.cuda.synchronize()
            secs = e0.elapsed_time(e1) * 1e-3
            flops = 8.0 * (1 << K) * (1 << n)            # 4 FMA = 8 flop per complex multiply-add
            per_gate[f"dense_k{K}"] = {"kernel": "k_dense_big<double,2048> (register-blocked)", "launches": 1,
                                       "ms": secs * 1e3, "tflops": flops / secs / 1e12,
                                       "frac_fp64": flops / secs / fp64_peak,
                                       "gbps": 2 * 16.0 * (1 << n) / secs / 1e9,
                                       "frac": 2 * 16.0 * (1 << n) / secs / 1e9 / peak,
                                       "bound": "fp64" if K >= 5 else "hbm"}
        roofline["per_gate_kernels"] = per_gate

        # Measurement-side kernels outside the expval of the step: the two sweeps of a native
        # mid-circuit measurement (single-wire marginal: S read; collapse: S/2 read + S written,
        # apply_operation.py:415-497) and a two-wire reduced density matrix (S read,
        # math/quantum.py:386-487).  Same timing rule: CUDA events around a few launches.
        S_bytes = 16.0 * 2 ** n
        meas = {}
        mprobes = {
            "mid_measure_probs": ("k_probs_marginal<double>", S_bytes,
                                  [lambda w=w: sv.probs_device([w]) for w in (0, n // 2)]),
            "reduced_dm_2_wires": ("k_gram_block<double,2,same>", S_bytes,
                                   [lambda: sv.reduced_dm([0, n // 2]), lambda: sv.reduced_dm([1, n - 1])]),
            "mid_measure_collapse": ("k_collapse<double>", 1.5 * S_bytes,
                                     [lambda w=w: sv.collapse(w, 0, False, 1.0) for w in (0, n // 2)]),
        }
        for name, (kern, byt, calls) in mprobes.items():
            for c in calls:
                c()
            torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for c in calls:
                c()
            e1.record()
            torch.cuda.synchronize()
            gb = byt * len(calls) / (e0.elapsed_time(e1) * 1e-3) / 1e9
            meas[name] = {"kernel": kern, "launches": len(calls), "gbps": gb, "frac": gb / peak}
        roofline["measurement_kernels"] = meas

    # Closed-form check at FULL size through the same fused path (nothing else pins correctness
    # above the ~24 qubits the oracle can hold): RY(theta_w) on every wire, then a CNOT chain
    # 0 -> 1 -> ... -> n-1: <Z_k> = prod_{j <= k} cos(theta_j).
    closed = None
    if args.fusion == "on" and not args.quick:
        th = np.random.default_rng(11).uniform(0.2, 1.2, n)
        cops = [q.RY(float(th[w]), wires=w) for w in range(n)] + \
               [q.CNOT(wires=[w, w + 1]) for w in range(n - 1)]
        sv.reset()
        sv.apply_operations_fused(cops, level=args.fusion_level)
        errs = {}
        for k in sorted({0, 1, n // 2, n - 2, n - 1}):
            got = float(measure(qb.expval(q.PauliZ(wires=k)), sv))
            errs[k] = abs(got - float(np.prod(np.cos(th[: k + 1]))))
        closed = {"circuit": f"RY(theta_w) on {n} wires + CNOT chain, <Z_k> = prod_(j<=k) cos(theta_j)",
                  "wires_checked": sorted(errs), "max_abs_err": max(errs.values()),
                  "norm2_minus_1": float(sv.norm2()) - 1.0}
        assert closed["max_abs_err"] < 1e-12, closed

    # e2e: public API, host parameters in / host scalar out, wall clock
    dev = qb.B200Qubit(wires=n, seed=0, fusion=args.fusion_level if args.fusion == "on" else 0)
    par = np.random.default_rng(3).uniform(0, 2 * np.pi, (args.layers, n, 2))
    pinned = torch.from_numpy(par).pin_memory()

    def e2e_step():
        host_par = pinned.numpy()
        ops2 = []
        for l in range(args.layers):
            for w in range(n):
                ops2.append(q.RY(float(host_par[l, w, 0]), wires=w))
                ops2.append(q.RZ(float(host_par[l, w, 1]), wires=w))
            for w in range(n):
                ops2.append(q.CNOT(wires=[w, (w + 1) % n]))
        t = qb.QuantumScript(ops2, [qb.expval(q.PauliZ(wires=0))])
        return dev.execute(t)

    e2e_step()
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r = e2e_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    assert abs(float(r) - float(val)) < 1e-9, (r, val)
    # per step: the coefficient tables (kernel parameters) and two tensor maps per launch go
    # host->device; without the specialised path every gate's matrix / phases do
    if use_program:
        h2d = sum(8 * int(p.ncoef) + 256 for p in prog.plans if p is not None)
    else:
        h2d = sum(64 if o.name == "RY" else 32 if o.name == "RZ" else 64 for o in ops_)
    e2e = {"value": ngates / e2e_s, "unit": "gates/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": 8, "seconds_per_step": e2e_s}

    # adjoint Jacobian s/step (the other half of BASELINE's metric)
    adjoint = None
    if not args.no_adjoint:
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res, jac = dev.execute_and_compute_derivatives(tape)      # first call: allocates ket + bra
        torch.cuda.synchronize()
        adj_first = time.perf_counter() - t0
        reps = 1 if args.quick else 2
        t0 = time.perf_counter()
        for _ in range(reps):
            res, jac = dev.execute_and_compute_derivatives(tape)
        torch.cuda.synchronize()
        adj_s = (time.perf_counter() - t0) / reps
        S = 16.0 * (1 << n)
        # what the reference's algorithm moves (one sweep per gate) ...
        adj_bytes = ngates * 2 * S + sum(2 * credited_bytes(o, n) for o in ops_)  # fwd + (ket+bra)
        # ... and what the fused reverse sweep moves: 2*S per forward segment, 4*S (ket + bra,
        # read + write) per reverse segment, 3*S to form the bra from the ket
        rev_segments = None
        fused_bytes = None
        if args.fusion == "on":
            from pennylane_b200.adjoint import _fused_reverse_program
            from pennylane_b200.compiler import merge_blocks, pack_segments
            T2, RB2, _ = sv.rt_geometry(2)
            _, L2 = sv.default_tile(2)
            rprog = _fused_reverse_program(tape, n, RB2, args.fusion_level)
            if rprog is not None:
                rev_segments = len(pack_segments(merge_blocks(rprog[0], args.fusion_level, fold_cx=not use_program),
                                                 n, T=T2, L=min(L2, T2), max_ops=64))
                fused_bytes = (len(fused_segments) * 2 + rev_segments * 4 + 3) * S
        adjoint = {"seconds_per_step": adj_s, "first_call_seconds": adj_first, "params": len(jac), "n_obs": 1,
                   "reverse_segments": rev_segments,
                   "fused_gbps": fused_bytes / adj_s / 1e9 if fused_bytes else None,
                   "frac_of_hbm_peak": fused_bytes / adj_s / 1e9 / peak if fused_bytes else None,
                   "per_gate_algorithmic_gbps": adj_bytes / adj_s / 1e9,
                   "grad_norm": float(np.linalg.norm(np.array(jac, dtype=float))),
                   # the step as one kernel family: 2*S per forward segment + 4*S (ket and bra,
                   # read + write) per reverse segment + 3*S to form the bra, over the WHOLE call
                   # (host work included), against the measured copy peak
                   "roofline": {"bound": "hbm", "kernel": "sk_kernel (NV = 2: ket + bra, generator inner "
                                "products fused)" if use_program else "k_rtile<double,3,2,512,1>",
                                "achieved": fused_bytes / adj_s / 1e9 if fused_bytes else None,
                                "peak": peak, "unit": "GB/s",
                                "frac": fused_bytes / adj_s / 1e9 / peak if fused_bytes else None,
                                "traffic": None}}
    clk = clocks.stop()

    cpu = None
    if not args.no_cpu_baseline:
        v, desc, cores, _, _ = cpu_baseline_sample(n, gates_per_kind=4 if args.quick else 8)
        cpu = {"value": v, "unit": "gates/s", "cores": cores, "kind": "port", "sample": desc}

    line = {
        "metric": "gates_per_s", "value": value, "unit": "gates/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "c128",
        "data": "synthetic", "config": workload_config(args, n),
        "hbm_gbps": roofline["all_gates_gbps"], "expval": float(val),
        "state_sweeps_per_step": (len(fused_segments) if fused_segments is not None else ngates),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "adjoint": adjoint,
        "closed_form_check": closed,
        "program_cache": dict(_program.STATS) if use_program else None,
        "gpu_launches": int(timed_launches), "clocks": clk,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=30, help="qubits per GPU shard (30 = 16 GiB)")
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--fusion", default="on", choices=["off", "on"])
    ap.add_argument("--fusion-level", type=int, default=1)
    ap.add_argument("--no-adjoint", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--weak16g", action="store_true",
                    help="multi-GPU: 30 + log2 N qubits (16 GiB per GPU) instead of 33 + log2 N (128 GiB)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
