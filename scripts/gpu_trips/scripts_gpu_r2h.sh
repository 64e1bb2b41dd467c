#!/bin/bash
# Round 2, GPU call H: epilogue rework (200 registers, single-chunk tiles straight from TMEM, uniform column-block lookup)
mkdir -p gpurun_out/r2h
O=gpurun_out/r2h
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5 > $O/pytest_tc.log; cat $O/pytest_tc.log
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_dualcnn.py tests/test_gpu_concnn.py -x -q 2>&1 | tail -8 > $O/pytest_parity.log; cat $O/pytest_parity.log | cut -c1-300
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers.json > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench2.log 2>&1; tail -1 $O/bench2.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['kernel_breakdown_ms_per_step'])"
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/tc_timing.log 2>&1; grep -c tc_timing $O/tc_timing.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision 3xtf32 > $O/bench_tf32.log 2>&1; tail -1 $O/bench_tf32.log | cut -c1-200
