#!/usr/bin/env python
"""Debug helper: the symmetric-memory exchange of ShardedStateVector at a given local size, with a
pulse per piece (torchrun, 2 ranks)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pennylane_b200.sharded import ExchangeStep, ShardedStateVector  # noqa: E402

nl = int(sys.argv[1]) if len(sys.argv) > 1 else 33
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{os.environ['LOCAL_RANK']}"))
t0 = time.time()


def log(msg):
    if rank == 0:
        print(f"[debug +{time.time() - t0:6.1f}s] {msg}", file=sys.stderr, flush=True)


sv = ShardedStateVector(nl + 1, dist, dtype=np.complex128, fusion=1)
log(f"state allocated: {torch.cuda.memory_allocated() >> 30} GiB")
data = sv.engine.data
ex = ExchangeStep([0], 1)
st = sv._symm_stage(1 << 26, data.dtype, data.device)
log(f"symm stage: {st is not None} {getattr(sv, '_symm_error', '')}")
torch.cuda.synchronize(); dist.barrier()
buf, hdl, cap = sv._symm
q = rank & 1
partners = [(1 - q, 1 - rank)]
chunk = 1 << (nl - 1)
piece = 1 << 26
row = data[0]
it = 0
for off in range(0, chunk, piece):
    pp = it & 1
    it += 1
    j, r = partners[0]
    buf[pp * cap: pp * cap + piece].copy_(row[j * chunk + off: j * chunk + off + piece])
    torch.cuda.synchronize(); log(f"piece {it}: staged")
    hdl.barrier(channel=pp)
    torch.cuda.synchronize(); log(f"piece {it}: barrier")
    remote = hdl.get_buffer(r, (piece,), data.dtype, pp * cap)
    row[j * chunk + off: j * chunk + off + piece].copy_(remote)
    torch.cuda.synchronize(); log(f"piece {it}: pulled")
    if it >= int(os.environ.get("MAXP", 6)):
        break
log("done")
dist.barrier()
dist.destroy_process_group()
