#!/bin/bash
# Two MMA-issuing warps on alternate tiles (HRNET_B200_PATCH_MMA2 / HRNET_B200_IGEMM_MMA2): parity, forward time, per-op table
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
HRNET_B200_PATCH_MMA2=1 HRNET_B200_IGEMM_MMA2=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/pytest_kernels_mma2.log 2>&1; echo "pytest kernels MMA2 rc=$?"; tail -4 gpurun_out/pytest_kernels_mma2.log
for v in "0 0" "1 0" "1 1" "1 96" "1 192" "0 1"; do set -- $v
echo "PATCH_MMA2=$1 IGEMM_MMA2=$2"; HRNET_B200_PATCH_MMA2=$1 HRNET_B200_IGEMM_MMA2=$2 timeout 200 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_s21.log
done
HRNET_B200_PATCH_MMA2=1 HRNET_B200_IGEMM_MMA2=1 timeout 200 python tools/op_roofline.py > gpurun_out/op_roofline_mma2.txt 2>&1; echo "op_roofline rc=$?"
head -24 gpurun_out/op_roofline_mma2.txt; tail -1 gpurun_out/op_roofline_mma2.txt
HRNET_B200_PATCH_MMA2=1 timeout 200 python tools/exp/l2_sweep.py b0 b1 2> gpurun_out/l2_sweep_mma2.log; grep "^\[dbg\]" gpurun_out/l2_sweep_mma2.log | awk 'NR%3==0' | cut -c1-420
