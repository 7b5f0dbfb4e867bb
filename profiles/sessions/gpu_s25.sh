#!/bin/bash
# Round 2, session 25: bench lines of the three configs with the current kernels + full GPU suite.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for cfg in w48 w32 poseresnet50; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 5 > gpurun_out/bench_$cfg.json 2> gpurun_out/bench_$cfg.err; echo "bench $cfg rc=$?"
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$cfg.json') if l.startswith('{')][-1])
print('$cfg', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', {k: d['roofline'].get(k) for k in ('achieved','frac','frac_of_sustained_peak','peak')}, 'launches', d.get('gpu_launches'), 'cpu', d['cpu_baseline']['value'], d['clocks'])
PY
done
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
