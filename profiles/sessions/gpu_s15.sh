#!/bin/bash
# Round 2, session 15: lean issue loops everywhere (per-conv im2col kernel too), 8 accumulators in the narrow patch chains; split sweep.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_kernels.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep -v "^stage2" gpurun_out/p.log; grep "stage4.0.branches" gpurun_out/p.err | grep -v "grid=148"; }
run debug nochain
run 400,250,175,175 380,250,185,185 380,230,195,195 400,230,185,185 420,250,165,165 400,270,165,165 390,240,170,200 390,240,200,170
run pair 400,250,175,175 380,250,185,185 380,230,195,195 400,230,185,185 420,250,165,165 400,270,165,165 390,240,170,200 390,240,200,170
