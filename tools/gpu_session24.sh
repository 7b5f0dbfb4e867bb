#!/bin/bash
cd "$(dirname "$0")/.."
K=3,5,16,33
run() { timeout 200 python tools/dbg_invariance.py "$1" $K 2>&1 | tail -1; }
run 0
HRNET_B200_NO_PDL=1 run 0
run 2
run 8
run 16
HRNET_B200_EPI=coal run 0
HRNET_B200_PATCH_MMA2=0 HRNET_B200_IGEMM_MMA2=0 HRNET_B200_NO_PDL=1 run 0
