#!/bin/bash
# Round 2, GPU call B: full device suite (after the C3 / gather fixes) + the reworked gather kernel
mkdir -p gpurun_out/r2b
O=gpurun_out/r2b
timeout 300 python scripts/bench_gather.py > $O/gather_2013.json 2> $O/gather_2013.err; grep -A4 '"v2_' $O/gather_2013.json | grep -E "v2_|ms|GB"; tail -3 $O/gather_2013.err
timeout 300 python bench.py --workload gather_c2 --steps 50 --warmup 5 > $O/bench_gather_c2.log 2>&1; tail -1 $O/bench_gather_c2.log | cut -c1-900
timeout 300 python bench.py --workload gather_c2 --batch 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_gather_c2_64k.log 2>&1; tail -1 $O/bench_gather_c2_64k.log | cut -c1-900
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_rows_kernel -c 1 -o $O/gather_rows -f \
   python bench.py --workload gather_c2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_gather.log 2>&1; tail -2 $O/ncu_gather.log
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
