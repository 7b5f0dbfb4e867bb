#!/bin/bash
# First GPU session of the next round (prepared at the end of round 1, when the GPU budget was spent):
#   1. refresh the judged evidence with the current kernels: tests, smoke, invariance sweep, bench, reference arm,
#      ncu launch list, ncu --set full of the stage-4 branch convs, per-op roofline  (= tools/gpu_final.sh)
#   2. epilogue cost micro-benchmark (what paces the C=48 convs once two issuers feed the tensor pipe)
#   3. role timers of the four branch convs with one / two issuers
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash tools/gpu_final.sh 2>&1 | tee gpurun_out/final.log
timeout 300 python tools/exp/run_epilogue_cost.py > gpurun_out/epilogue_cost.log 2>&1; echo "epilogue_cost rc=$?"; tail -40 gpurun_out/epilogue_cost.log
for m in 0 1; do
  HRNET_B200_PATCH_MMA2=$m timeout 200 python tools/exp/l2_sweep.py b0 b1 2> gpurun_out/roles_patch_mma2_$m.log
  grep "^\[dbg\]" gpurun_out/roles_patch_mma2_$m.log | awk 'NR%3==0' | cut -c1-420
done
