#!/bin/bash
# 2-GPU bench sweep over environment settings
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  i=$((i+1))
  env $cfg timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-2} --master-addr 127.0.0.1 --master-port $((29530+i)) bench.py --gpus ${NG:-2} --steps 3 --warmup 2 > gpurun_out/sweep2_$i.json 2> gpurun_out/sweep2_$i.err; echo "[$cfg] rc=$?"
  python -c "import sys,json; d=json.loads(open('gpurun_out/sweep2_$i.json').read().strip().split('\n')[-1]); print('   ', round(d['value']), 'ms', round(d['ms_per_step']), 'sweeps', d['state_sweeps_per_step'], 'frac', round(d['roofline']['frac'],3), 'run share', round(d['roofline']['share_of_step'],3), 'exch s', round(d['exchange']['seconds_per_step'],3))" || tail -3 gpurun_out/sweep2_$i.err
done
