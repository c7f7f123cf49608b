#!/usr/bin/env python
"""Quick GPU micro-benchmark: total fused forward time of the 30-qubit ansatz + the empty-program
copy rate, for the current environment knobs (B200Q_RT_*)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from pennylane_b200.compiler import Segment, compile_ops
from pennylane_b200.statevector import StateVector

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
L = int(os.environ.get("B200Q_TILE_L", 5))
sv = StateVector(n)
T = sv.rt_geometry(1)[0]
segs = compile_ops(bench.hea_ops(n), n, level=1, T=T, L=L)
for s in segs:
    sv.run_segment(s)
torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
sv.reset(); a.record()
for s in segs:
    sv.run_segment(s)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b)
seg = Segment(list(range(L)) + list(range(n - (T - L), n)), [], 0)
sv.run_segment(seg); torch.cuda.synchronize()
a.record()
for _ in range(5):
    sv.run_segment(seg)
b.record(); torch.cuda.synchronize()
print(json.dumps({"env": {k: v for k, v in os.environ.items() if k.startswith("B200Q_")},
                  "segments": len(segs), "total_ms": round(ms, 1), "gates_per_s": round(720 / ms * 1e3),
                  "empty_gbps": round(2 * 16 * (1 << n) * 5 / (a.elapsed_time(b) * 1e-3) / 1e9)}))

if "--adjoint" in sys.argv:
    import time
    import pennylane_b200 as qb
    del sv
    torch.cuda.empty_cache()
    tape = bench.hea_tape(n)
    dev = qb.B200Qubit(wires=n, fusion=1)
    for i in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res, jac = dev.execute_and_compute_derivatives(tape)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps({"adjoint_s": round(dt, 3), "grad_norm": float((sum(float(j) ** 2 for j in jac)) ** 0.5)}))
