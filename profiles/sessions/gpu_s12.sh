#!/bin/bash
# Round 2, session 12: what slows the MMA stream of the im2col chain?  Epilogue stripped / TMEM loads only / global only.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for v in skip0 skip1 skip2 skip3 "pair skip0" "pair skip1"; do
  echo "=== $v"
  timeout 300 python tools/chain_probe.py debug $v 2>&1 | grep "im2col\|C=192\|C=384\|forward"
done
