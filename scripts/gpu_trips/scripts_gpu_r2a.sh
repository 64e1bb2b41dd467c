#!/bin/bash
# Round 2, GPU call A (one B200): full device suite, gather v1 vs vector kernel (+ncu), headline bench with the
# target-list e2e, C3 bench, whole-scene inference, compute-sanitizer over the GEMM building-block tests.
mkdir -p gpurun_out/r2a
O=gpurun_out/r2a
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/pytest_gpu.log; cat $O/pytest_gpu.log
timeout 300 python scripts/bench_gather.py > $O/gather_2013.json 2> $O/gather_2013.err; tail -40 $O/gather_2013.json; tail -3 $O/gather_2013.err
timeout 300 python scripts/bench_gather.py --grss2018 > $O/gather_2018.json 2> $O/gather_2018.err; tail -30 $O/gather_2018.json; tail -3 $O/gather_2018.err
timeout 300 python bench.py --workload gather_c2 --steps 50 --warmup 5 > $O/bench_gather_c2.log 2>&1; tail -1 $O/bench_gather_c2.log
timeout 300 python bench.py --workload gather_c2 --batch 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_gather_c2_64k.log 2>&1; tail -1 $O/bench_gather_c2_64k.log
timeout 300 python bench.py --workload gather_c3 --steps 50 --warmup 5 > $O/bench_gather_c3.log 2>&1; tail -1 $O/bench_gather_c3.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log
timeout 600 python bench.py --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51.log 2>&1; tail -1 $O/bench_c3_51.log
timeout 600 python scripts/bench_inference.py > $O/inference.json 2> $O/inference.err; tail -1 $O/inference.json; tail -3 $O/inference.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_rows_kernel -c 2 -o $O/gather_rows -f \
   python bench.py --workload gather_c2 --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_gather.log 2>&1; tail -2 $O/ncu_gather.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q > $O/sanitizer_memcheck_tc.log 2>&1; tail -5 $O/sanitizer_memcheck_tc.log
