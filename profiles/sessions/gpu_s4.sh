#!/bin/bash
# Round 2, session 4: proxy fence moved out of the producers (it serialised the operand loads).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_reference_seam.py -x -q --timeout=300 -p no:cacheprovider 2>&1 | tail -8
echo "=== probe"
timeout 900 python tools/chain_probe.py debug nochain "400,200,200,200" "330,240,215,215" "305,250,228,217" "275,260,237,228" "250,270,240,240" > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log
grep "chain-dbg" gpurun_out/chain_probe.err | grep -E "stage2.0|stage4.0" | cut -c1-420
tail -3 gpurun_out/chain_probe.err
echo "=== full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
