#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
HRNET_B200_DBG=1 timeout 300 python tools/profile_convs.py 64 2 2>&1 | tee gpurun_out/dbg_convs.log | grep dbg | awk 'NR%2==0'
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches_serial.csv python tools/profile_forward.py 64 10 > gpurun_out/prof_forward.log 2>&1; echo "ncu launches rc=$?"
