#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/exp/clock_probe.py > gpurun_out/clock_probe.log 2>&1; echo "clock_probe rc=$?"; head -30 gpurun_out/clock_probe.log
timeout 300 python tools/exp/cold_sweep.py 2> gpurun_out/cold_sweep.log; echo "cold_sweep rc=$?"
grep -c "dbg-ns" gpurun_out/cold_sweep.log
