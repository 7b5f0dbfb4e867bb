#!/bin/bash
# Round 2, session 26: SM split of the branch chains by in-situ module times.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python tools/split_sweep.py 0,0,0,0 407,251,171,171 392,243,182,183 380,250,185,185 380,260,180,180 365,265,185,185 400,260,170,170 420,250,165,165 392,258,175,175 375,245,190,190 410,240,175,175 2>&1 | grep -v Warning | tee gpurun_out/split_sweep.log
timeout 900 python tools/split_sweep.py pair 0,0,0,0 407,251,171,171 392,243,182,183 420,250,165,165 430,260,155,155 2>&1 | grep -v Warning | tee gpurun_out/split_sweep_pair.log
