#!/bin/bash
# Round 2, session 2: chain scheduler v2 (chunked tickets, one acquire per chunk) + four epilogue warpgroups.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=300 -p no:cacheprovider 2>&1 | tail -15
echo "=== probe"
timeout 600 python tools/chain_probe.py debug nochain "367,216,209,208" "316,278,203,203" "300,300,200,200" "280,320,200,200" "250,250,250,250" > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log
grep "chain-dbg" gpurun_out/chain_probe.err | grep -E "stage2.0|stage3.0|stage4.0" | cut -c1-330
tail -3 gpurun_out/chain_probe.err
