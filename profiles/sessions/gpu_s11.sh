#!/bin/bash
# Round 2, session 11: ncu --set full of the stage-4.0 chain kernels (single CTA / CTA pairs) -- what bounds the MMA rate?
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
prof() { timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 \
   -o gpurun_out/prof_$4 -f python tools/profile_forward.py 64 0 $5 > gpurun_out/prof_$4.log 2>&1; echo "ncu $4 rc=$?"; }
prof "conv_chain_igemm" 4 2 s11_igemm "20=0"
prof "conv_chain_igemm" 4 2 s11_igemm_pair "20=2"
prof "conv_chain_patch" 10 2 s11_patch "20=0"
ls -la gpurun_out/*.ncu-rep
