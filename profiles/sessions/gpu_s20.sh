#!/bin/bash
# Round 2, session 20: which part of the epilogue slows the C = 48 MMAs: TMEM loads or global traffic?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches.[01]" gpurun_out/p.err | grep -v "grid=148"; }
run debug skip2
run debug skip3
run debug skip1
