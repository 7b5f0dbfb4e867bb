#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('PARTITIONED', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], [(b['C'],b['us_avg']) for b in d['roofline']['per_branch']])"
tail -3 gpurun_out/bench.err
