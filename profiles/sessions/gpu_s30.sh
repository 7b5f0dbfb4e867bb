#!/bin/bash
# Round 2, session 30: device-side multi-person crops (Pillow bilinear, bit for bit), head + fused argmax (one pixel per thread).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; grep "^conv1 \|final_layer\|argmax\|serial total" gpurun_out/op_roofline.txt
