#!/bin/bash
# Round 2, session 1: first contact of the branch-chain kernels (hang-safe: everything under `timeout`).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 300 python -m pytest tests/test_gpu_chain.py -x -q --timeout=200 -p no:cacheprovider -k "plan_is_active or per_conv" 2>&1 | tail -25
echo "=== chain full size"
timeout 400 python -m pytest tests/test_gpu_chain.py -x -q --timeout=300 -p no:cacheprovider -k "full_size or two_forwards" 2>&1 | tail -25
echo "=== variants"
timeout 300 python tools/variants_bench.py 0 128 2>&1 | tail -4
echo "=== full gpu suite"
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; head -14 gpurun_out/op_roofline.txt; tail -1 gpurun_out/op_roofline.txt
