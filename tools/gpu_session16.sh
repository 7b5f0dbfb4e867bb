#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HRNET_B200_EPI=$EPI timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -4 gpurun_out/pytest_kernels.log
run() { echo "EPI=$1 CS=$2 MINK=$3"; HRNET_B200_EPI=$1 HRNET_B200_CS=$2 HRNET_B200_CS_MINK=$3 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_unroll.log; }
run auto 1 0
run direct 1 0
run auto 2 1000
run auto 2 0
HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 2>&1 | grep "^\[dbg\]\|^shape" | awk '/^shape/ {print; n=0} /^\[dbg\]/ {n++; if (n==2) print}' | cut -c1-460 > gpurun_out/dbg_unroll.log; cat gpurun_out/dbg_unroll.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; head -50 gpurun_out/op_roofline.txt
