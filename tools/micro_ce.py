#!/usr/bin/env python
"""Copy-engine micro-benchmark for the exchange (torchrun, 2 GPUs): NVLink push and local unpack
of 1 GiB as contiguous / pitched 2D copies (b200q_remap_copy), alone and beside an HBM-saturating
kernel on the compute stream.  Prints GB/s per variant."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pennylane_b200._lib import check, load  # noqa: E402


def main():
    rank = int(os.environ["RANK"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    import torch.distributed._symmetric_memory as symm
    lib = load()
    GiB = 1 << 30
    big_gib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    state = torch.empty(big_gib * GiB // 8, dtype=torch.float64, device="cuda")
    state.zero_()
    stage = symm.empty(2 * GiB // 8, dtype=torch.float64, device="cuda")
    hdl = symm.rendezvous(stage, dist.group.WORLD)
    peer = hdl.buffer_ptrs[1 - rank]
    load_buf = torch.empty(8 * GiB // 8, dtype=torch.float64, device="cuda")
    load_buf.zero_()
    comm = torch.cuda.Stream()
    out = {}

    def copy(dst, dp, src, sp, run, count, stream):
        check(lib.b200q_remap_copy(C.c_void_p(dst), dp, C.c_void_p(src), sp, run, count, C.c_void_p(stream)))

    variants = {}
    for name, run, pitch in (("contig", GiB, GiB), ("rows16M_pitch128M", 16 << 20, 128 << 20),
                             ("rows1M_pitch8M", 1 << 20, 8 << 20), ("rows128K_pitch1M", 128 << 10, 1 << 20)):
        count = GiB // run
        if pitch * count > big_gib * GiB:
            continue
        variants["push_" + name] = (peer, run, state.data_ptr(), pitch, run, count)
        variants["unpack_" + name] = (state.data_ptr(), pitch, stage.data_ptr(), run, run, count)
        variants["pack_" + name] = (stage.data_ptr() + GiB, run, state.data_ptr(), pitch, run, count)
    for loaded in (False, True):
        for name, (dst, dp, src, sp, run, count) in variants.items():
            torch.cuda.synchronize(); dist.barrier()
            reps = 8
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            if loaded:
                for _ in range(40):
                    load_buf.mul_(1.0000001)          # 16 GiB of traffic per call at ~6 TB/s: ~2.7 ms
            with torch.cuda.stream(comm):
                e0.record(comm)
                for _ in range(reps):
                    copy(dst, dp, src, sp, run, count, comm.cuda_stream)
                e1.record(comm)
            torch.cuda.synchronize()
            out[("loaded_" if loaded else "alone_") + name] = round(reps * GiB / (e0.elapsed_time(e1) * 1e-3) / 1e9, 1)
    if rank == 0:
        print(json.dumps(out, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
