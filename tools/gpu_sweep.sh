#!/bin/bash
# bench sweep: each argument is a string of environment assignments (plus optional bench args after --)
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs="$cfg"; args=""
  if [[ "$cfg" == *" -- "* ]]; then args="${cfg#* -- }"; envs="${cfg%% -- *}"; fi
  env $envs timeout 900 python bench.py --no-cpu-baseline $args > gpurun_out/sweep_$i.json 2> gpurun_out/sweep_$i.err; echo "[$cfg] rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/sweep_$i.json').read().strip().split('\n')[-1]); print('   ', round(d['value']), round(d['roofline']['frac'],3), d['state_sweeps_per_step'], 'e2e', round(d['e2e']['value']), 'adjoint', (d.get('adjoint') or {}).get('seconds_per_step'))" || tail -3 gpurun_out/sweep_$i.err
done
