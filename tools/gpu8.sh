#!/bin/bash
# 8-GPU session: weak-scaling bench (33 q), 36-qubit bench, 36-qubit C5 sampling, world-4 tests
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
tail -1 gpurun_out/bench_8gpu.json | cut -c1-300
timeout 200 $TR --master-port 29522 tools/run_c5_sharded.py --qubits 36 --fast-sampling > gpurun_out/c5_8gpu_36q.json 2> gpurun_out/c5_8gpu.err; echo "c5 rc=$?"
cat gpurun_out/c5_8gpu_36q.json
timeout 240 $TR --master-port 29523 bench.py --gpus 8 --qubits 33 --steps 2 --warmup 1 > gpurun_out/bench_8gpu_36q.json 2> gpurun_out/bench_8gpu_36q.err; echo "bench8-36q rc=$?"
tail -1 gpurun_out/bench_8gpu_36q.json | cut -c1-300

nvidia-smi --query-gpu=memory.total --format=csv | head -2
