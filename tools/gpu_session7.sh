#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
HRNET_B200_DBG=1 timeout 300 python tools/profile_convs.py 64 2 2>&1 | grep dbg | tee gpurun_out/dbg_convs.log | awk 'NR>8' | grep -v "dbg-ns" | cut -c1-520
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=300 -p no:cacheprovider -x > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_forward.log 2>&1; echo "pytest forward rc=$?"; tail -4 gpurun_out/pytest_forward.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('PAIR', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], [(b['C'],b['us_avg']) for b in d['roofline']['per_branch']])"
tail -3 gpurun_out/bench.err
HRNET_B200_CS=1 timeout 300 python tools/variants_bench.py 0
timeout 300 python tools/variants_bench.py 0
