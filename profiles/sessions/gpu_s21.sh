#!/bin/bash
# Round 2, session 21: full GPU suite after the issue-loop / constant-bank / fuse index changes; split sweep; pairs.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split\|launch" gpurun_out/p.log; }
run debug nochain 407,251,171,171 390,250,180,180 380,250,185,185 370,250,190,190 380,240,190,190 380,260,180,180 380,250,175,195 380,250,195,175
run pair 407,251,171,171 390,250,180,180 380,250,185,185 400,250,175,175 420,250,165,165
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; head -50 gpurun_out/op_roofline.txt; tail -1 gpurun_out/op_roofline.txt
