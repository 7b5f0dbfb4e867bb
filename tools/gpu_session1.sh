#!/bin/bash
# First GPU session: diagnostics, then the pytest GPU suite. Everything is bounded by `timeout`.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/diag.log
tail -40 gpurun_out/diag.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
