#!/bin/bash
# Round 2, session 23: two forwards in flight (two plans, two streams) vs one at a time; split sweep with the robust timer.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 400 python tools/chain_probe.py "$@" > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward\|split\|launch" gpurun_out/p.log; tail -2 gpurun_out/p.err; }
run dual 407,251,171,171 395,255,175,175 380,260,180,180 380,250,185,185 365,265,185,185 395,240,182,183
run pair dual 407,251,171,171 380,260,180,180
