#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider -x > gpurun_out/pytest_kernels.log 2>&1; echo "pytest kernels rc=$?"; tail -6 gpurun_out/pytest_kernels.log
HRNET_B200_BPS=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider -x > gpurun_out/pytest_kernels_bps1.log 2>&1; echo "pytest kernels bps1 rc=$?"; tail -3 gpurun_out/pytest_kernels_bps1.log
timeout 900 python -m pytest tests/test_gpu_forward.py -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_forward.log 2>&1; echo "pytest forward rc=$?"; tail -6 gpurun_out/pytest_forward.log
run() { echo "BPS=$1 CS=$2 MINK=$3"; HRNET_B200_BPS=$1 HRNET_B200_CS=$2 HRNET_B200_CS_MINK=$3 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_issue.log; }
run 1 1 0
run "" 1 0
run "" 2 1000
run "" 2 0
HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 2>&1 | grep "^\[dbg\]\|^shape" | awk '/^shape/ {print; n=0} /^\[dbg\]/ {n++; if (n==2) print}' | cut -c1-460 > gpurun_out/dbg_issue.log; cat gpurun_out/dbg_issue.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; head -48 gpurun_out/op_roofline.txt
