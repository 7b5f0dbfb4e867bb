#!/bin/bash
# Round 2, session 5: deferred tile publication + deep residual prefetch; one vs two M-tiles per im2col ticket; device resize.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 600 python -m pytest tests/test_gpu_chain.py tests/test_gpu_kernels.py tests/test_gpu_forward.py -x -q --timeout=300 -p no:cacheprovider -k "chain or cubic or device_resize" 2>&1 | tail -8
echo "=== probe (two M-tiles per im2col ticket)"
timeout 900 python tools/chain_probe.py debug "305,250,228,217" "290,240,240,230" "270,240,250,240" > gpurun_out/chain_probe.log 2> gpurun_out/chain_probe.err
cat gpurun_out/chain_probe.log
grep "chain-dbg" gpurun_out/chain_probe.err | grep -E "stage2.0|stage4.0" | cut -c1-420
tail -2 gpurun_out/chain_probe.err
echo "=== probe (one M-tile per im2col ticket)"
timeout 900 python tools/chain_probe.py m1 debug nochain "305,250,228,217" "330,260,210,200" "350,270,190,190" > gpurun_out/chain_probe_m1.log 2> gpurun_out/chain_probe_m1.err
cat gpurun_out/chain_probe_m1.log
grep "chain-dbg" gpurun_out/chain_probe_m1.err | grep -E "stage4.0" | cut -c1-420
tail -2 gpurun_out/chain_probe_m1.err
