#!/bin/bash
# Round 2, session 31: head + fused argmax with the redux.sync warp stage.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; grep "^conv1 \|final_layer\|argmax\|serial total" gpurun_out/op_roofline.txt
