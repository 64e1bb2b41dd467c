#!/bin/bash
# GPU trip: tensor-core building block first (short timeout: a hang must not eat the box), then parity, bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc.py -m gpu -x -q --no-header --tb=short 2>&1 | tail -30 > gpurun_out/pytest_tc.log
tail -15 gpurun_out/pytest_tc.log
if ! grep -q "passed" gpurun_out/pytest_tc.log || grep -q "failed\|error" gpurun_out/pytest_tc.log; then echo "TC TESTS NOT GREEN"; exit 1; fi
timeout 900 python -m pytest tests -m gpu -q --no-header --tb=short 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --prof-out gpurun_out/prof_layers.json > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log | cut -c1-400
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 2> gpurun_out/tc_timing.txt; tail -59 gpurun_out/tc_timing.txt | cut -c1-330 | grep -v "image_gen_net_[123]\|fc_final\|fc_2\|fc_1"
