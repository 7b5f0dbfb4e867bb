#!/bin/bash
# Round 2, session 17: lean issuer in the halo-patch chains; one vs two issuers; splits.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -3
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches.[01]" gpurun_out/p.err | grep -v "grid=148"; }
run debug 400,250,175,175 370,250,190,190 340,250,205,205
run debug t13=1 400,250,175,175 370,250,190,190 340,250,205,205
run debug skip9 400,250,175,175
