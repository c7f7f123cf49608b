#!/bin/bash
# bench sweep over kernel variant / run length (forward only)
mkdir -p gpurun_out
show() { python -c "import sys,json; d=json.loads(open('$1').read().strip().split('\n')[-1]); print('$2', round(d['value']), round(d['roofline']['frac'],3), d['state_sweeps_per_step'], d.get('adjoint',{}).get('seconds_per_step'))"; }
for cfg in ${CFGS:-"1 4" "1 5" "0 5" "0 4"}; do
  set -- $cfg
  B200Q_RT_VARIANT=$1 B200Q_TILE_L=$2 timeout 600 python bench.py --no-cpu-baseline --no-adjoint > gpurun_out/sweep_v$1_L$2.json 2> gpurun_out/sweep_v$1_L$2.err; echo "rc=$?"
  show gpurun_out/sweep_v$1_L$2.json "v$1 L$2"
done
