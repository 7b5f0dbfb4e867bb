#!/bin/bash
# Round 2, session 18: halo-patch chains without epilogue work (what bounds the C = 48 chain at 2,150 clk per tile?)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split" gpurun_out/p.log; grep "stage4.0.branches.[01]" gpurun_out/p.err; }
run debug skip1
run debug skip1 t13=1
run debug t4=20
