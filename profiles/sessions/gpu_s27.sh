#!/bin/bash
# Round 2, session 27: chains of a module starting individually (as soon as their own input is summed) vs together; A B A B.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for i in 1 2; do
  timeout 300 python tools/chain_probe.py nochain 2>/dev/null | grep "default"
  timeout 300 python tools/chain_probe.py t23=1 nochain 2>/dev/null | grep "tune\|default"
done
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=400 -p no:cacheprovider 2>&1 | tail -2
