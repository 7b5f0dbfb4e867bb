#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider -k "pair" > gpurun_out/pytest_pair.log 2>&1; echo "pytest pair rc=$?"; tail -15 gpurun_out/pytest_pair.log
for v in "0 9999" "16 9999" "16 64" "80 128" "16 128"; do set -- $v; echo "PATCH_PAIR=$1 MAX=$2"; HRNET_B200_PATCH_PAIR=$1 HRNET_B200_PATCH_PAIR_MAX=$2 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_pair.log; done
HRNET_B200_PATCH_PAIR=16 timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline_pair.txt 2>&1; head -12 gpurun_out/op_roofline_pair.txt
