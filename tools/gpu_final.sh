#!/bin/bash
# Round-end verification + evidence session: GPU tests, smoke, invariance sweep, bench (+ CPU baseline), reference arms,
# ncu launch list of one forward, ncu --set full of the kernels the roofline talks about, racecheck / memcheck of the smoke run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/dbg_invariance.py 0 3,5,8,16,33,64 2>&1 | tail -1   # forward(x[:k]) == forward(x)[:k], bit for bit, three runs each
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('BENCH', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], [(b['C'],b['us_per_conv']) for b in d['roofline']['per_branch']], 'cpu', d['cpu_baseline'], d['clocks']); print({k:(v['us'],v['ops']) for k,v in d['layer_breakdown']['classes'].items()})"
tail -3 gpurun_out/bench.err
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py 64 > gpurun_out/prof_forward.log 2>&1; echo "ncu launches rc=$?"
# ncu --set full of the stage-4.0 kernels + the HBM-bound ends.  A forward has 16 halo-patch chains (stage4.0 = launches
# 10-11), 10 im2col chains (stage4.0 = 4-5), 20 exchange units (stage4.0 = 12-15), 23 sums (stage4.0 = 14-17).
prof() { timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"$1" -s $2 -c $3 \
   -o gpurun_out/prof_$4 -f python tools/profile_forward.py 64 > gpurun_out/prof_$4.log 2>&1; echo "ncu $4 rc=$?"; }
prof "conv_chain_patch" 10 2 chain_patch
prof "conv_chain_igemm" 4 2 chain_igemm
prof "conv_xunit" 12 4 xunit
prof "fuse_sum" 14 4 fuse
prof "head_c_kernel|head_argmax_finish|stem_conv3x3s2_tc" 0 3 ends
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
