#!/bin/bash
# Round 2, session 13: deeper operand pipelines in the im2col chains (single CTA: 5 stages, CTA pairs: 7-8).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep -v "^stage2\|C= 48\|C= 96" gpurun_out/p.log; grep "stage4.0.branches" gpurun_out/p.err | grep -v "grid=148"; }
run debug stages4
run debug
run debug skip1
run pair debug stages4
run pair debug 367,216,209,208 400,250,175,175
run pair debug skip1
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -3
