#!/bin/bash
# One GPU session: smoke -> gpu tests -> bench (prefetch off/on) -> segment microbench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests.log
tail -5 gpurun_out/tests.log
B200Q_RT_PREFETCH=0 timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pf0.json 2> gpurun_out/bench_pf0.err; echo "bench pf0 rc=$?"
cat gpurun_out/bench_pf0.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('pf0', d['value'], d['roofline']['frac'], d.get('adjoint'))"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_pf1.json 2> gpurun_out/bench_pf1.err; echo "bench pf1 rc=$?"
cat gpurun_out/bench_pf1.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('pf1', d['value'], d['roofline']['frac'], d.get('adjoint'))"
timeout 600 python tools/microbench_rtile.py 30 > gpurun_out/micro_rtile.json 2> gpurun_out/micro_rtile.err; echo "micro rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/micro_rtile.json').read().strip().split('\n')[-1])
for k,v in d.items():
    if isinstance(v,dict): print(k, v['segments'], round(v['total_ms'],1), round(v['gates_per_s']), round(v['gbps']))
    else: print(k, round(v))
PY
