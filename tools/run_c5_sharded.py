#!/usr/bin/env python
"""BASELINE config 5 on N GPUs (torchrun): random circuit of depth 20 on n qubits (36 on 8 B200s:
128 GiB of complex128 per GPU) + 1 000 000 shots, statevector sharded by the high-order qubits.

  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_c5_sharded.py --qubits 36

Prints one JSON line on rank 0: circuit seconds, sampling seconds, exchange volume, checks
(sample shape, ones fraction ~ 0.5 for a scrambling circuit, norm).  Not part of bench.py."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import pennylane_b200 as qb
    from pennylane_b200 import ops as q
    from pennylane_b200.sharded import ShardedStateVector

    ap = argparse.ArgumentParser()
    ap.add_argument("--qubits", type=int, default=36)
    ap.add_argument("--depth", type=int, default=20)
    ap.add_argument("--shots", type=int, default=1000000)
    ap.add_argument("--fast-sampling", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.qubits
    rng = np.random.default_rng(5)
    ops_ = []
    for _ in range(a.depth):
        for i in range(n):
            ops_.append(q.Rot(*rng.uniform(0, 2 * np.pi, 3), wires=i))
        perm = rng.permutation(n)
        for x, y in zip(perm[::2], perm[1::2]):
            ops_.append(q.CNOT(wires=[int(x), int(y)]))
    sv = ShardedStateVector(n, dist, dtype=np.complex128, fusion=1)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    program = sv.compile(ops_)
    t_compile = time.perf_counter() - t0
    sv.reset()
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    sv.run(program)
    torch.cuda.synchronize(); dist.barrier()
    t_run = time.perf_counter() - t0
    norm = float(sv.norm2())
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    s = sv.sample(a.shots, np.random.default_rng(5), None, not a.fast_sampling)
    torch.cuda.synchronize(); dist.barrier()
    t_sample = time.perf_counter() - t0
    if rank == 0:
        s = np.asarray(s)
        print(json.dumps({
            "config": f"c5: {n}q random circuit depth {a.depth} ({len(ops_)} gates) + {a.shots} shots on {world} GPUs, complex128",
            "n_gpus": world, "shard_bytes": int(16 * (1 << (n - (world.bit_length() - 1)))),
            "compile_seconds": t_compile, "circuit_seconds": t_run, "gates_per_s": len(ops_) / t_run,
            "sampling_seconds": t_sample, "sampling_mode": "fast (blocked scan)" if a.fast_sampling else "exact (numpy addition order)",
            "exchange_bytes_sent_per_gpu": int(sv.stats.get("exchange_bytes", 0)),
            "norm2": norm, "sample_shape": list(s.shape), "ones_fraction": float(s.mean())}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
