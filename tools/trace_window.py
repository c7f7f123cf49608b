#!/usr/bin/env python
"""Timeline of the overlapped exchange windows (torchrun, >= 2 GPUs): per piece, when the
segments before the exchange ran, when the communication stream packed / waited at the barrier /
pulled, and when the segments after the exchange ran.  Times in ms from the start of the step.
  torchrun --nproc-per-node 2 tools/trace_window.py [qubits per GPU = 30] [layers = 3]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200.sharded import ShardedStateVector  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
    g = world.bit_length() - 1
    nl = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    n = nl + g
    sv = ShardedStateVector(n, dist, dtype=np.complex128, fusion=1)
    prog = sv.compile(bench.hea_ops(n, layers))
    for _ in range(2):
        sv.reset(); sv.run(prog)
    torch.cuda.synchronize(); dist.barrier()
    sv.trace = []
    t0 = torch.cuda.Event(enable_timing=True)
    sv.reset()
    t0.record()
    sv.run(prog)
    torch.cuda.synchronize(); dist.barrier()
    if rank >= 0:
        sched = [(e[0], len(e[1]) if e[0] == "window" else None, len(e[3]) if e[0] == "window" else None,
                  e[4:] if e[0] == "window" else None) for e in prog["schedule"]]
        rows = []
        for tag, p, evs in sv.trace:
            rows.append((tag, p, [round(t0.elapsed_time(e), 2) for e in evs]))
        out = json.dumps({"n": n, "rank": rank, "schedule": sched, "trace": rows})
        tag = os.environ.get("TRACE_TAG", "trace")
        os.makedirs("gpurun_out", exist_ok=True)
        with open(f"gpurun_out/{tag}_rank{rank}.json", "w") as f:
            f.write(out)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
