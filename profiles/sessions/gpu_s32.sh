#!/bin/bash
# Round 2, session 32: fuse_sum_kernel with block-uniform row math (the kernel is issue-bound, not DRAM-bound).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_chain.py -q --timeout=600 -p no:cacheprovider 2>&1 | tail -2
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; grep "fuse\.[0-9]\|serial total" gpurun_out/op_roofline.txt
timeout 300 python tools/chain_probe.py dual 2>/dev/null | grep "one forward"
