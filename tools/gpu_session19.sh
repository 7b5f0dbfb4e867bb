#!/bin/bash
# Diagnostics: what bounds the C=192 / C=384 branch convs (grid-cap / CTA-pair / M2 sweeps with role timers), pair mode
# for those convs only in the whole forward, per-op table in that mode.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/exp/l2_sweep.py 2> gpurun_out/l2_sweep.log; echo "l2_sweep rc=$?"
grep -c "dbg-ns" gpurun_out/l2_sweep.log
echo "default"; timeout 200 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_s19.log
for mk in 1700 3000; do
echo "CS_MINK=$mk"; HRNET_B200_CS=2 HRNET_B200_CS_MINK=$mk timeout 200 python tools/variants_bench.py 0 2>&1 | tee -a gpurun_out/variants_s19.log
done
HRNET_B200_CS=2 HRNET_B200_CS_MINK=1700 timeout 200 python tools/op_roofline.py > gpurun_out/op_roofline_pair1700.txt 2>&1; echo "op_roofline rc=$?"
head -14 gpurun_out/op_roofline_pair1700.txt
