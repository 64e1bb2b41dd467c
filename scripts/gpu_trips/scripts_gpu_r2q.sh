#!/bin/bash
# Round 2, GPU call Q: gather with shuffle addressing (tests + bench), ncu --set full of a bn_apply launch
mkdir -p gpurun_out/r2q gpurun_out/ncu
O=gpurun_out/r2q
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "gather" 2>&1 | tail -4 > $O/pytest_gather.log; cat $O/pytest_gather.log | cut -c1-300
timeout 300 python scripts/bench_gather.py --only v2 > $O/gather_2013.json 2> $O/gather_2013.err; grep -A3 '"v2_' $O/gather_2013.json | grep -E "v2_|GB|\"ms"; tail -2 $O/gather_2013.err
timeout 300 python scripts/bench_gather.py --only v2 --grss2018 > $O/gather_2018.json 2> $O/gather_2018.err; grep -A3 '"v2_' $O/gather_2018.json | grep -E "v2_|GB"
timeout 300 python bench.py --workload gather_c2 --steps 50 --warmup 5 > $O/bench_gather_c2.log 2>&1; tail -1 $O/bench_gather_c2.log | cut -c1-500
timeout 300 python bench.py --workload gather_c2 --batch 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_gather_c2_64k.log 2>&1; tail -1 $O/bench_gather_c2_64k.log | cut -c1-500
timeout 300 python bench.py --workload gather_c3 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_gather_c3.log 2>&1; tail -1 $O/bench_gather_c3.log | cut -c1-500
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_bn_apply_kernel -s 24 -c 1 -o $O/bn_apply -f python scripts/one_step.py --steps 2 > $O/ncu_bn_apply.log 2>&1; tail -1 $O/ncu_bn_apply.log
ncu -i $O/bn_apply.ncu-rep --page raw --csv > $O/bn_apply.raw.csv 2>/dev/null; ncu -i $O/bn_apply.ncu-rep --page source --csv > $O/bn_apply.source.csv 2>/dev/null; rm -f $O/bn_apply.ncu-rep
