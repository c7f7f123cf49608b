#!/bin/bash
# 2-GPU session: sharded GPU tests (windows included), then bench lines at the given settings.
#   TAG=r2g bash tools/gpu2_r2.sh "ENV=.. ARGS" ...   each argument: "VAR=val VAR=val -- bench args"
TAG=${TAG:-r2g}
NG=${NG:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_segjit.py -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "tests rc=$?"
  tail -5 gpurun_out/${TAG}_tests.log
fi
i=0
for cfg in "$@"; do
  i=$((i+1))
  envs="${cfg%%--*}"; bargs="${cfg#*--}"
  env $envs timeout ${BENCH_TIMEOUT:-500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 \
    --master-port $((29530+i)) bench.py --gpus $NG $bargs > gpurun_out/${TAG}_$i.json 2> gpurun_out/${TAG}_$i.err; echo "[$cfg] rc=$?"
  python - <<PY || tail -5 gpurun_out/${TAG}_$i.err
import json
d = json.loads(open('gpurun_out/${TAG}_$i.json').read().strip().split('\n')[-1])
x = d['exchange']
print('   ', round(d['value'], 1), 'gates/s  ms', round(d['ms_per_step']), 'sweeps', d['state_sweeps_per_step'],
      'seg frac', round(d['roofline']['frac'] or 0, 3), 'amp/s/gpu', '%.3g' % d['amplitude_updates_per_s_per_gpu'],
      '| exch windows', x.get('overlapped_with_sweeps'), 'plain', x.get('not_overlapped'),
      'comm s', x.get('comm_stream_seconds_per_step'), 'GB/s', round(x['sent_gbps_per_gpu'] or 0),
      'visible s', round(x['visible_seconds_per_step'], 3), 'share', round(x['share_of_step'], 3),
      '| parity', d['parity']['max_abs_err_state'], d['parity']['samples_identical'])
PY
done
