#!/bin/bash
# Round 2, session 8 (2 GPUs): batch-sharded bench with the pipelined all-gather and the bit-exact gather check.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-kernel-roofline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -3 gpurun_out/bench_2gpu.err
python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu.json')); print('2GPU', d['value'], d['ms_per_step'], 'per-gpu', d['step_flops']['per_gpu_persons_per_s'], 'e2e', d['e2e']['value'], 'gather_check', d.get('gather_check'), d['clocks'])"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-kernel-roofline --no-cpu-baseline > gpurun_out/bench_1gpu_same_box.json 2>> gpurun_out/bench_2gpu.err; python -c "
import json; d=json.load(open('gpurun_out/bench_1gpu_same_box.json')); print('1GPU', d['value'], d['ms_per_step'])"
