#!/bin/bash
# Round 2, session 34 (2 GPUs): sharded bench, gather_check.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
grep "^{" gpurun_out/bench_2gpu.log | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('2GPU', d['value'], d['ms_per_step'], d.get('ms_per_step_per_rank'), d.get('gather_check'), 'e2e', d['e2e']['value'])"
tail -3 gpurun_out/bench_2gpu.err
