#!/bin/bash
# Round 2, session 9: exchange sums as tickets of the exchange-unit kernel.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=300 -p no:cacheprovider 2>&1 | tail -12
echo "=== probe"
timeout 900 python tools/chain_probe.py nochain > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log; tail -3 gpurun_out/chain_probe.err
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; head -30 gpurun_out/op_roofline.txt; tail -1 gpurun_out/op_roofline.txt
echo "=== full gpu suite"
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
