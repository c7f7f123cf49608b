"""Build ``pennylane_b200/csrc/libb200q.so`` in-tree with nvcc for sm_100a.

``python -m pennylane_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(CSRC, "libb200q.so")
SOURCES = ["api.cu", "dm.cu", "rtile.cu", "segk_host.cu", "remap.cu"] + [f"rtile_k_{p}_{k}.cu" for p in "df"
                                    for k in ("fwd", "ws", "adj", "adj2")]          # compiled in parallel, one object each
HEADERS = ["common.cuh", "gates.cuh", "measure.cuh", "sample.cuh", "adjoint.cuh", "tile.cuh",
           "rtile.cuh", "rtile_host.h", "rtile_launch.cuh", "segk_args.h", os.path.join("..", "..", "include", "b200q.h")]
OBJDIR = os.path.join(CSRC, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libb200q.so")


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    for f in SOURCES + HEADERS:
        p = os.path.join(CSRC, f)
        if os.path.exists(p) and os.path.getmtime(p) > t:
            return True
    return False


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(OBJDIR, os.path.splitext(src)[0] + ".o")
        srcp = os.path.join(CSRC, src)
        deps = [srcp] + [os.path.join(CSRC, h) for h in HEADERS]
        if not force and os.path.exists(obj) and all(
                os.path.getmtime(d) <= os.path.getmtime(obj) for d in deps if os.path.exists(d)) \
                and src not in _always_rebuild(obj):
            return obj, ""
        cmd = [nvcc, *NVCC_FLAGS, "-c", srcp, "-o", obj]
        if verbose:
            cmd[1:1] = ["-Xptxas", "-v"]
        res = subprocess.run(cmd, capture_output=True, text=True, cwd=CSRC)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
        return obj, res.stderr

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as pool:
        results = list(pool.map(compile_one, SOURCES))
    if verbose:
        for _, log in results:
            print(log)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-Xcompiler", "-fPIC",
           *[o for o, _ in results], "-ldl", "-o", OUT]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=CSRC)
    if res.returncode != 0:
        raise RuntimeError(f"link failed:\n{' '.join(cmd)}\n{res.stdout}\n{res.stderr}")
    return OUT


def _always_rebuild(obj):
    return ()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
