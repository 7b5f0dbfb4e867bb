#!/bin/bash
# Round 2, session 38: early barrier probes in the three-slot issuer of the C = 96 chains.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q --timeout=600 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/chain_probe.py debug > gpurun_out/p.log 2> gpurun_out/p.err; grep "forward" gpurun_out/p.log; grep "stage4.0.branches.1" gpurun_out/p.err | grep -v "grid=148"
timeout 300 python tools/split_sweep.py 0,0,0,0 385,215,200,200 2>&1 | grep -v Warning
