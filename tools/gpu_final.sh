#!/bin/bash
# Full verification + measurement session (round end): GPU tests, smoke, bench (+ CPU baseline), reference arm,
# ncu launch list of one forward, ncu --set full of the four stage-4 branch convs, memcheck of the smoke run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out; rm -f gpurun_out/parity.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python tools/dbg_invariance.py 0 3,5,8,16,33,64 2>&1 | tail -1   # forward(x[:k]) == forward(x)[:k], bit for bit, three runs each
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('BENCH', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roof', d['roofline']['achieved'], d['roofline']['frac'], [(b['C'],b['us_avg']) for b in d['roofline']['per_branch']], 'cpu', d['cpu_baseline'], d['clocks']); print({k:(v['us'],v['launches']) for k,v in d['layer_breakdown']['classes'].items()})"
tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"; cat gpurun_out/bench_reference.json | cut -c1-400
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/launches.csv python tools/profile_forward.py 64 > gpurun_out/prof_forward.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_igemm|conv3x3_patch" -o gpurun_out/prof_convs -f \
   python tools/profile_convs.py > gpurun_out/prof_convs.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/op_roofline.py > gpurun_out/op_roofline.txt 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
