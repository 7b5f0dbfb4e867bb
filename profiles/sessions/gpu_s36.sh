#!/bin/bash
# Round 2, session 36: full GPU suite + bench with the three-slot C = 96 chains and the re-calibrated SM split.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('BENCH', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], 'traffic', d['roofline']['traffic'], d['clocks'])"
timeout 300 python tools/split_sweep.py 0,0,0,0 2>&1 | grep -v Warning
