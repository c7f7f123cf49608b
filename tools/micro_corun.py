#!/usr/bin/env python
"""What can run BESIDE a fused segment launch?  The segment kernel is persistent (2 CTAs per SM
for the whole launch, 5.8 ms at 30 qubits).  While a queue of segment launches runs on the
compute stream, a second (high-priority) stream issues one operation at a time and we time it:
contiguous / pitched device-to-device copies (b200q_remap_copy), a tiny torch kernel, stream
memory operations.  An operation that needs an SM slot it cannot get waits for a launch boundary."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from pennylane_b200._lib import check, load  # noqa: E402
from pennylane_b200.statevector import StateVector  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
lib = load()
sv = StateVector(n)
segs = sv.compile_fused(bench.hea_ops(n, 2), level=1)
sv.prepare_segments(segs)
seg = segs[int(os.environ.get("SEG", "1"))]
GiB = 1 << 30
MODE = int(os.environ.get('MODE', '0'))
src = torch.zeros(GiB // 8 * 2, dtype=torch.float64, device="cuda")
dst = torch.zeros(GiB // 8 * 9, dtype=torch.float64, device="cuda")
flag = torch.zeros(16, dtype=torch.int32, device="cuda")
comm = torch.cuda.Stream(priority=-1)
cs = comm.cuda_stream


def copy(d, dp, s, sp, run, count):
    check(lib.b200q_remap_copy(C.c_void_p(d), dp, C.c_void_p(s), sp, run, count, C.c_void_p(cs)))


ops = {
    "copy_1d_1GiB": lambda: copy(dst.data_ptr(), GiB, src.data_ptr(), GiB, GiB, 1),
    "copy_2d_rows16M": lambda: copy(dst.data_ptr(), 128 << 20, src.data_ptr(), 16 << 20, 16 << 20, 64),
    "copy_2d_rows1M": lambda: copy(dst.data_ptr(), 8 << 20, src.data_ptr(), 1 << 20, 1 << 20, 1024),
    "copy_1d_x64_rows16M": lambda: [copy(dst.data_ptr() + i * (128 << 20), 16 << 20, src.data_ptr() + i * (16 << 20),
                                         16 << 20, 16 << 20, 1) for i in range(64)],
    "unpack_kernel_1d_1GiB": lambda: check(lib.b200q_remap_unpack(C.c_void_p(dst.data_ptr()), GiB, C.c_void_p(src.data_ptr()), GiB, GiB, 1, MODE, 0, C.c_void_p(cs))),
    "unpack_kernel_rows1M": lambda: check(lib.b200q_remap_unpack(C.c_void_p(dst.data_ptr()), 8 << 20, C.c_void_p(src.data_ptr()), 1 << 20, 1 << 20, 1024, MODE, 0, C.c_void_p(cs))),
    "unpack_kernel_rows1M_296ctas": lambda: check(lib.b200q_remap_unpack(C.c_void_p(dst.data_ptr()), 8 << 20, C.c_void_p(src.data_ptr()), 1 << 20, 1 << 20, 1024, MODE, 296, C.c_void_p(cs))),
    "tiny_kernel": lambda: flag.add_(1),
    "memop_write": lambda: check(lib.b200q_stream_write32(C.c_void_p(flag.data_ptr() + 32), 7, C.c_void_p(cs))),
}
out = {}
for loaded in (False, True):
    for name, fn in ops.items():
        ts = []
        for rep in range(3):
            torch.cuda.synchronize()
            if loaded:
                for _ in range(6):
                    sv.run_segment(seg)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(comm):
                e0.record(comm)
                fn()
                e1.record(comm)
            torch.cuda.synchronize()
            ts.append(round(e0.elapsed_time(e1), 3))
        out[("beside_segments_" if loaded else "alone_") + name + "_ms"] = ts
# segment time for reference
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(6):
    sv.run_segment(seg)
e1.record()
torch.cuda.synchronize()
out["segment_ms"] = round(e0.elapsed_time(e1) / 6, 3)
# correctness of the kernel copy
src.copy_(torch.arange(src.numel(), dtype=torch.float64, device="cuda"))
dst.zero_()
check(lib.b200q_remap_unpack(C.c_void_p(dst.data_ptr()), 8 << 20, C.c_void_p(src.data_ptr()), 1 << 20, 1 << 20, 1024, MODE, 0,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))
torch.cuda.synchronize()
want = src[: GiB // 8].view(1024, -1)
got = torch.as_strided(dst, (1024, (1 << 20) // 8), ((8 << 20) // 8, 1))
out["unpack_kernel_correct"] = bool(torch.equal(want, got))
plan = seg._sk_plan
out["segment_rounds"] = len(plan.rounds)
print(json.dumps(out))
