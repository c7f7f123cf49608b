#!/bin/bash
# One GPU session: smoke -> gpu tests -> bench (variants) -> segment microbench.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/tests.log
tail -5 gpurun_out/tests.log
show() { python -c "import sys,json; d=json.loads(open('$1').read().strip().split('\n')[-1]); print('$2', round(d['value']), round(d['roofline']['frac'],3), d['state_sweeps_per_step'], d.get('adjoint',{}).get('seconds_per_step'))"; }
for v in ${VARIANTS:-0 1}; do
  B200Q_RT_VARIANT=$v timeout 600 python bench.py --no-cpu-baseline $BENCH_ARGS > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err; echo "bench v$v rc=$?"
  show gpurun_out/bench_v$v.json v$v
done
timeout 600 python tools/microbench_rtile.py 30 > gpurun_out/micro_rtile.json 2> gpurun_out/micro_rtile.err; echo "micro rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/micro_rtile.json').read().strip().split('\n')[-1])
for k,v in d.items():
    if isinstance(v,dict):
        print(k, v['segments'], round(v['total_ms'],1), round(v['gates_per_s']), round(v['gbps']))
        print('   ', v['per_segment'][:8])
    else: print(k, round(v))
PY
# mid-circuit measurement / reduced-density-matrix kernels and the one-shot loop (profiles/r1g_mcm.json)
timeout 400 python tools/bench_mcm.py ${MCM_ARGS:-30 26 400} > gpurun_out/mcm.json 2> gpurun_out/mcm.err; echo "mcm rc=$?"
python - <<'PY'
import json
d = json.load(open('gpurun_out/mcm.json'))
for k, v in d['kernels'].items():
    if isinstance(v, dict):
        print(k, round(v['ms'], 3), round(v['frac'], 3))
print(d['one_shot'])
PY
