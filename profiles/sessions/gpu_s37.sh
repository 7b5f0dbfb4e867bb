#!/bin/bash
# Round 2, session 37: ncu --set full of the final halo-patch chains (three-slot C = 96) -- refreshes the evidence.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"conv_chain_patch" -s 10 -c 2 \
   -o gpurun_out/prof_chain_patch -f python tools/profile_forward.py 64 > gpurun_out/prof_chain_patch.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py 64 > gpurun_out/prof_forward.log 2>&1; echo "ncu launches rc=$?"
