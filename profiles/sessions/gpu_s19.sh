#!/bin/bash
# Round 2, session 19: BN constants of the chain epilogues from the kernel parameters (LDC) instead of shared memory / L1.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_forward.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches" gpurun_out/p.err | grep -v "grid=148"; }
run debug 400,250,175,175 420,250,165,165 380,250,185,185 400,230,185,185 420,230,175,175 440,240,160,160
run debug t13=1 400,250,175,175
run pair 400,250,175,175 420,250,165,165 380,250,185,185
