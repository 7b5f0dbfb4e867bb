#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for e in auto tma; do
HRNET_B200_EPI=$e timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout=120 -p no:cacheprovider > gpurun_out/pytest_kernels_$e.log 2>&1; echo "pytest kernels EPI=$e rc=$?"; tail -4 gpurun_out/pytest_kernels_$e.log
done
for v in "0 0" "16 64" "16 128" "16 48"; do set -- $v
for ep in auto; do
echo "PAIR=$1 MAX=$2 EPI=$ep"; HRNET_B200_EPI=$ep HRNET_B200_PATCH_PAIR=$1 HRNET_B200_PATCH_PAIR_MAX=$2 timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_wtma.log
done; done
echo "PAIR=0 EPI=tma"; HRNET_B200_EPI=tma timeout 300 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_wtma.log
for pp in 0 16; do for ep in direct tma; do echo "PAIR=$pp EPI=$ep"; HRNET_B200_EPI=$ep HRNET_B200_PATCH_PAIR=$pp HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 64,96,72,48,48,3,1,1,2 64,96,72,48,48,3,1,0,2 2>&1 | grep "^\[dbg" | awk 'NR%4==3 || NR%4==0' | cut -c1-460 ; done; done > gpurun_out/dbg_wtma.log 2>&1; cat gpurun_out/dbg_wtma.log
