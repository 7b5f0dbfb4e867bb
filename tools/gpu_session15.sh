#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
for e in auto direct coal tma; do
HRNET_B200_EPI=$e timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/pytest_kernels_$e.log 2>&1; echo "pytest kernels EPI=$e rc=$?"; tail -4 gpurun_out/pytest_kernels_$e.log
done
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_forward.log 2>&1; echo "pytest forward rc=$?"; tail -6 gpurun_out/pytest_forward.log
run() { echo "EPI=$1 PAIR=$2 MAX=$3"; HRNET_B200_EPI=$1 HRNET_B200_PATCH_PAIR=$2 HRNET_B200_PATCH_PAIR_MAX=$3 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_coal.log; }
run direct 0 9999
run auto 0 9999
run coal 0 9999
run auto 16 9999
run auto 16 64
run auto 80 128
for pp in 0 16; do echo "PAIR=$pp EPI=auto"; HRNET_B200_PATCH_PAIR=$pp HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 64,96,72,48,48,3,1,1,2 64,48,36,96,96,3,1,1,2 64,24,18,192,192,3,1,1,1 2>&1 | grep "^\[dbg" | awk 'NR%4==3 || NR%4==0' | cut -c1-460 ; done > gpurun_out/dbg_coal.log 2>&1; cat gpurun_out/dbg_coal.log
