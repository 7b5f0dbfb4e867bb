#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -5 gpurun_out/pytest_kernels.log
run() { echo "PAIR=$1 MAX=$2 EPI=$3 NACC=$4"; HRNET_B200_PATCH_PAIR=$1 HRNET_B200_PATCH_PAIR_MAX=$2 HRNET_B200_EPI_TMA=$3 HRNET_B200_PATCH_NACC=$4 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_pair2.log; }
run 0 9999 3 2
run 0 9999 3 4
run 0 9999 2 4
run 16 9999 3 4
run 16 9999 2 4
run 16 64 2 4
run 80 128 2 4
for pp in 0 16; do echo "PAIR=$pp EPI=2"; HRNET_B200_EPI_TMA=2 HRNET_B200_PATCH_PAIR=$pp HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 64,96,72,48,48,3,1,1,2 64,48,36,96,96,3,1,1,2 2>&1 | grep "^\[dbg\]" ; done > gpurun_out/dbg_pair2.log 2>&1; cut -c1-420 gpurun_out/dbg_pair2.log
