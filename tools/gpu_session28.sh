#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 900 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log | cut -c1-200
timeout 200 python tools/dbg_invariance.py 0 3,5,8,16,33,64 2>&1 | tail -1 | cut -c1-400
timeout 200 python tools/variants_bench.py 0 2>&1 | tail -1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
