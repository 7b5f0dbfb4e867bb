#!/bin/bash
# Round 2, session 29: head with four pixels per thread (constant-bank weights) + fused argmax.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_forward.py tests/test_gpu_kernels.py -q -x --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1; grep "^conv1 \|final_layer\|argmax\|serial total" gpurun_out/op_roofline.txt
timeout 600 python bench.py --config w48 --steps 20 --warmup 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('BENCH', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])"
