#!/bin/bash
# Round 2, session 14: lean MMA issue loop in the im2col chains (single CTA and CTA pairs) vs the general loop.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep -v "^stage2\|C= 48\|C= 96" gpurun_out/p.log; grep "stage4.0.branches" gpurun_out/p.err | grep -v "grid=148"; }
run debug skip9
run debug
run debug stages4
run pair debug 367,216,209,208 400,250,175,175 430,270,150,150
run 367,216,209,208 400,250,175,175 430,270,150,150 460,290,125,125
