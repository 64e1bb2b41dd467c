#!/bin/bash
# Round 2, GPU call D: whole device suite (no -x: every failure in one trip), gather after the software pipeline,
# 3xF16 per-layer timings + ncu launch list
mkdir -p gpurun_out/r2d gpurun_out/ncu
O=gpurun_out/r2d
timeout 300 python bench.py --workload gather_c2 --steps 50 --warmup 5 > $O/bench_gather_c2.log 2>&1; tail -1 $O/bench_gather_c2.log | cut -c1-700
timeout 300 python bench.py --workload gather_c2 --batch 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_gather_c2_64k.log 2>&1; tail -1 $O/bench_gather_c2_64k.log | cut -c1-700
timeout 300 python scripts/bench_gather.py --only v2 > $O/gather_2013.json 2> $O/gather_2013.err; grep -A3 '"v2_' $O/gather_2013.json | grep -E "v2_|GB"
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers_3xf16.json > $O/bench_layers.log 2>&1; tail -1 $O/bench_layers.log | cut -c1-300
bash scripts/ncu_step.sh; cp gpurun_out/ncu/launches.csv $O/launches_3xf16.csv
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
