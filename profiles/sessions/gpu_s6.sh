#!/bin/bash
# Round 2, session 6: full verification of the current state + bench lines of the three configs + reference arms.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_w48.json 2> gpurun_out/bench_w48.err; echo "bench w48 rc=$?"; tail -2 gpurun_out/bench_w48.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_w48.json'))
print('W48', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['sync_call_value'], 'launches', d['launches_per_forward'])
print('roof', d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['us_total'], [(b['C'],b['kernel'],b['us_per_conv']) for b in d['roofline']['per_branch']])
print({k:(v['us'],v['ops']) for k,v in d['layer_breakdown']['classes'].items()})
print('cpu', d['cpu_baseline']); print(d['clocks'])
PY
for c in w32 poseresnet50; do timeout 900 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "bench $c rc=$?"; tail -2 gpurun_out/bench_$c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_$c.json')); print('$c', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'cpu', d['cpu_baseline']['value'], d['cpu_baseline']['kind'], d['cpu_baseline']['cores'])"; done
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/bench_reference.json
timeout 900 python bench.py --impl reference-cuda --steps 5 --warmup 2 > gpurun_out/bench_reference_cuda.json 2> gpurun_out/bench_reference_cuda.err; echo "refcuda rc=$?"; cut -c1-900 gpurun_out/bench_reference_cuda.json; tail -3 gpurun_out/bench_reference_cuda.err
