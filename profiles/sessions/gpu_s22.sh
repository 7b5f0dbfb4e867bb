#!/bin/bash
# Round 2, session 22: exchange units per source branch (lean issue loop, BN constants as kernel parameters).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_chain.py tests/test_gpu_forward.py -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split\|launch" gpurun_out/p.log; }
run nochain 407,251,171,171 380,260,180,180 380,250,185,185 370,250,190,190
run pair nochain 407,251,171,171 380,260,180,180
