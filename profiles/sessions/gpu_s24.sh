#!/bin/bash
# Round 2, session 24: mbarrier waits with a suspend-time hint (A) vs re-issued probes (B), sustained forward timing, A B A B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python tools/chain_probe.py dual 2>/dev/null | grep "one forward"
  timeout 300 python tools/chain_probe.py dual lib=simple-hrnet_b200/libhrnet_b200_nohint.so 2>/dev/null | grep "library\|one forward"
done
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -2
