#!/bin/bash
# Round 2, session 33: final bench lines of configs 2 and 5, the reference arms (CPU, cuDNN on the same GPU).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in w32 poseresnet50; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$cfg.json') if l.startswith('{')][-1])
print('$cfg', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', {k: d['roofline'].get(k) for k in ('achieved','frac')}, 'launches', d.get('gpu_launches'), 'cpu', d['cpu_baseline']['value'])
PY
done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; tail -1 gpurun_out/bench_reference.json | cut -c1-400
timeout 600 python bench.py --impl reference-cuda --steps 5 --warmup 2 > gpurun_out/bench_reference_cuda.json 2> gpurun_out/bench_reference_cuda.err; echo "reference-cuda rc=$?"; tail -1 gpurun_out/bench_reference_cuda.json | cut -c1-400
