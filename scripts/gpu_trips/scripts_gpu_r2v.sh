#!/bin/bash
# Round 2, last single-GPU trip: after the level-wgrad tap groups -- refreshes what changed in gpurun_out/r2z (device
# suite + smoke, bench lines, per-kernel profile, role timing, ncu launch list, C3, inference, sanitizer over a train step)
mkdir -p gpurun_out/r2z gpurun_out/ncu
O=gpurun_out/r2z
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest_gpu.log; grep -E "passed|failed|FAILED|ERROR" $O/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --precision 3xtf32 > $O/bench_tf32.log 2>&1; tail -1 $O/bench_tf32.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 > $O/bench_bf16.log 2>&1; tail -1 $O/bench_bf16.log | cut -c1-200
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers.json > $O/bench_layers.log 2>&1
timeout 600 python bench.py --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51.log 2>&1; tail -1 $O/bench_c3_51.log | cut -c1-300
timeout 600 python scripts/bench_inference.py > $O/inference.json 2> $O/inference.err; tail -1 $O/inference.json | cut -c1-300
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/tc_timing.log 2>&1
bash scripts/ncu_step.sh; cp gpurun_out/ncu/launches.csv $O/launches_3xf16.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/one_step.py --steps 1 --batch 256 > $O/sanitizer_memcheck_step.log 2>&1; tail -3 $O/sanitizer_memcheck_step.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/one_step.py --steps 1 --batch 128 > $O/sanitizer_racecheck_step.log 2>&1; tail -3 $O/sanitizer_racecheck_step.log
grep "and Read access\|and Write access" $O/sanitizer_racecheck_step.log | grep -v "hyp_tc.cuh:4[0-9][0-9]" | head -3
