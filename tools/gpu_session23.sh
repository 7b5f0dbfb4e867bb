#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in "0 0" "1 0" "0 96"; do set -- $v
echo "PATCH_MMA2=$1 IGEMM_MMA2=$2"
HRNET_B200_PATCH_MMA2=$1 HRNET_B200_IGEMM_MMA2=$2 timeout 300 python tools/dbg_invariance.py 2>&1 | tail -18
done
