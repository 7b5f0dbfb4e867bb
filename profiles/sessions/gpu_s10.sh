#!/bin/bash
# Round 2, session 10: im2col chains on CTA pairs (cta_group::2) -- first contact.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -12
echo "=== probe (default)"
timeout 600 python tools/chain_probe.py debug nochain > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log; tail -3 gpurun_out/chain_probe.err
echo "=== probe (pairs)"
timeout 600 python tools/chain_probe.py pair debug 367,216,209,208 380,240,190,190 400,250,175,175 420,260,160,160 > gpurun_out/chain_probe_pair.log 2> gpurun_out/chain_probe_pair.err
cat gpurun_out/chain_probe_pair.log; tail -3 gpurun_out/chain_probe_pair.err
