"""CPU: the sharded schedule of the benchmark circuit with the REAL lowering (fusion pass,
specialised-kernel plans) on every rank of a world, emulated in one process: the piece bits of
every overlapped exchange agree on all ranks (the collective OR of the busy masks in
ShardedStateVector._schedule), every segment of every rank yields a valid plan, and every
distinct structure compiles with NVRTC (which needs no GPU)."""
import numpy as np
import pytest
import torch

import bench
from pennylane_b200 import compiler as cc
from pennylane_b200 import segjit as sj
from pennylane_b200 import sharded as sh
from pennylane_b200._lib import load


class _Dist:
    """all_gather over ranks that run one after another: values are recorded per call and
    replayed; the caller iterates until the record is stable."""

    record: dict = {}

    def __init__(self, world, rank):
        self.w, self.r, self.calls = world, rank, 0

    def get_world_size(self, group=None):
        return self.w

    def get_rank(self, group=None):
        return self.r

    def all_gather(self, out, mine, group=None):
        idx = self.calls
        self.calls += 1
        _Dist.record.setdefault(idx, {})[self.r] = mine.clone()
        for k, o in enumerate(out):
            o.copy_(_Dist.record[idx].get(k, mine))


class _Engine:
    batch = 1
    data = None

    def __init__(self, nl, geom, L):
        self.nl, self.geom, self.L = nl, geom, L

    def reset(self, index=None):
        pass

    def compile(self, lops):
        g = self.geom
        return ("segs", cc.compile_ops(lops, self.nl, level=1, T=g.T, L=self.L, fold_cx=False, RB=g.RB, sww=g.sww))

    def units(self, handle):
        out = []
        for seg in handle[1]:
            ok = seg.tile_bits is not None and len(seg.tile_bits) == self.geom.T
            out.append((seg, sum(1 << b for b in seg.tile_bits) if ok else None))
        return out


@pytest.mark.parametrize("n,world", [(21, 2), (22, 4), (23, 8)])
def test_windows_agree_across_ranks_and_plans_compile(n, world):
    geom = sj.default_geometry(1, 1)
    L = sj.default_low_bits(1, 1)
    g = world.bit_length() - 1
    ops_ = bench.hea_ops(n, 4)
    _Dist.record = {}
    progs = {}
    for _ in range(5):
        before = {k: {r: v.clone() for r, v in d.items()} for k, d in _Dist.record.items()}
        for r in range(world):
            sv = sh.ShardedStateVector(n, _Dist(world, r), engine=_Engine(n - g, geom, L))
            progs[r] = sv.compile(ops_)
        if before.keys() == _Dist.record.keys() and all(
                torch.equal(before[k][r], _Dist.record[k][r]) for k in before for r in before[k]):
            break
    else:
        pytest.fail("the emulated all_gather did not reach a fixed point")

    def windows(prog):
        return [(e[0], e[4], e[5]) if e[0] == "window" else (e[0],)
                for e in prog["schedule"] if e[0] in ("window", "exchange")]

    ref = windows(progs[0])
    assert any(w[0] == "window" for w in ref)
    for r in range(world):
        assert windows(progs[r]) == ref, f"rank {r} chose other piece bits"
    plans = {}
    for r in range(world):
        for kind, item in progs[r]["steps"]:
            if kind != "run" or item is None:
                continue
            for seg in item[1]:
                if seg.tile_bits is None or len(seg.tile_bits) != geom.T:
                    continue
                p = sj.plan_segment(seg, geom, cc.low_run(seg.tile_bits))
                plans[p.key] = p
        # window piece bits are outside the tiles of the segments they pipeline
        for e in progs[r]["schedule"]:
            if e[0] == "window":
                mask = ((1 << e[5]) - 1) << e[4]
                for seg in list(e[1]) + list(e[3]):
                    assert not (mask & sum(1 << b for b in seg.tile_bits))
    assert plans
    if load().b200q_jit_available():
        sj.ensure_compiled(list(plans.values()))
