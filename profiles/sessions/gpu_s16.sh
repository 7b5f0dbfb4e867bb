#!/bin/bash
# Round 2, session 16: one MMA issuer vs two in the C = 48 halo-patch chain (8 accumulators), 4 vs 8 accumulators.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches.[01]" gpurun_out/p.err | grep -v "grid=148"; }
run debug 400,250,175,175
run debug t13=1 400,250,175,175
run debug t14=4 400,250,175,175
