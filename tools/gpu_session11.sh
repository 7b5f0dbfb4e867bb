#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -3 gpurun_out/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_forward.log 2>&1; echo "pytest forward rc=$?"; tail -6 gpurun_out/pytest_forward.log
for m in 0 2; do echo "EPI_TMA=$m"; HRNET_B200_EPI_TMA=$m timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_epi.log; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('BENCH', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], [(b['C'],b['us_avg']) for b in d['roofline']['per_branch']]); print({k:(v['us'],v['launches']) for k,v in d['layer_breakdown']['classes'].items()})"
tail -3 gpurun_out/bench.err
