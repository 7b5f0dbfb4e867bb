#!/bin/bash
# Round 2, session 35: C = 96 halo-patch chains with three patch slots (two wide + one 64-byte-row slot for the 32-channel chunk).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_forward.py -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches.1" gpurun_out/p.err | grep -v "grid=148"; }
run debug
run debug skip8
timeout 600 python tools/split_sweep.py 0,0,0,0 365,245,195,195 375,235,195,195 2>&1 | grep -v Warning
