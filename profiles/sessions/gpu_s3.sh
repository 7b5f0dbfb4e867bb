#!/bin/bash
# Round 2, session 3: row tickets + ring coordinates + 256-bit epilogue accesses; tune[] API instead of getenv.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_chain.py -x -q --timeout=300 -p no:cacheprovider 2>&1 | tail -8
echo "=== probe"
timeout 600 python tools/chain_probe.py debug nochain "367,216,209,208" "330,250,210,210" "300,280,210,210" "400,200,200,200" > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log
grep "chain-dbg" gpurun_out/chain_probe.err | grep -E "stage2.0|stage4.0" | cut -c1-420
tail -3 gpurun_out/chain_probe.err
echo "=== full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
