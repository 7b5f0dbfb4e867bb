#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for pp in 0 16; do for ns in "" 1; do for res in 1 0; do
echo "PAIR=$pp NOSTORE=$ns res=$res"
env HRNET_B200_PATCH_PAIR=$pp ${ns:+HRNET_B200_DBG_NOSTORE=1} HRNET_B200_DBG=1 timeout 120 python tools/dbg_shapes.py 64,96,72,48,48,3,1,$res,2 64,48,36,96,96,3,1,$res,2 2>&1 | grep "^\[dbg" | awk 'NR%4==3 || NR%4==0' | cut -c1-460
done; done; done > gpurun_out/dbg_pair3.log 2>&1
cat gpurun_out/dbg_pair3.log
