#!/bin/bash
# Round 2, GPU call T3: store path modes (0 plain blocks, 1 TMA everywhere, 2 TMA for the dgrad reductions), per-kernel times
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
for rep in 1 2; do
for d in 0 1 2; do
HYP_TC_TMA_STORE=$d timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_mode$d.log 2>&1; tail -1 $O/bench_mode$d.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('mode $d rep $rep', round(d['ms_per_step'],3), d['clocks'])"
done
done
for d in 0 1 2; do
HYP_TC_TMA_STORE=$d HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_mode$d.json > $O/prof_mode$d.log 2>&1
done
