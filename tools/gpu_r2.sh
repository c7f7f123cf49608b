#!/bin/bash
# One GPU session of round 2: smoke -> gpu tests -> bench -> ncu launch list of the bench command
# -> ncu --set full of the specialised segment kernel (forward + reverse sweep).
# Outputs under gpurun_out/<TAG>_*; the summaries kept for the judge are copied to profiles/.
TAG=${TAG:-r2f}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ -z "$SKIP_TESTS" ]; then
  timeout 300 python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/${TAG}_smoke.log
  tail -3 gpurun_out/${TAG}_smoke.log
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
  tail -5 gpurun_out/${TAG}_tests.log
fi
timeout 900 python bench.py $BENCH_ARGS > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().split('\n')[-1])
print('bench', round(d['value']), 'gates/s frac', round(d['roofline']['frac'], 3), 'e2e', round(d['e2e']['value']),
      'adjoint', d.get('adjoint', {}).get('seconds_per_step'))
print({k: (round(v.get('frac', 0), 3), v.get('frac_per_wires'), v.get('frac_fp64')) for k, v in d['roofline']['per_gate_kernels'].items()})
PY
if [ -z "$SKIP_NCU" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-adjoint \
    > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sk_kernel -c 4 \
    -o gpurun_out/${TAG}_segk_fwd30 -f python tools/prof_segk.py 30 4,3 > gpurun_out/${TAG}_ncu_fwd.log 2>&1; echo "ncu fwd rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:sk_kernel -s 32 -c 6 \
    -o gpurun_out/${TAG}_segk_adj28 -f python tools/prof_segk_adj.py 28 > gpurun_out/${TAG}_ncu_adj.log 2>&1; echo "ncu adj rc=$?"
fi
