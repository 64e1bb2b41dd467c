#!/bin/bash
# First GPU call of round 2 (one B200): everything round 1 left unmeasured, in order of value.
#   gpurun --timeout 900 -- 'bash scripts_gpu_round2.sh'
mkdir -p gpurun_out
# 1. the whole device suite (the apps / loaders / samplers added late in round 1 ran only in their own files)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
# 2. patch gather: v1 vs v2 (bit-compared inside the script), GB/s against the measured copy peak
timeout 300 python scripts/bench_gather.py > gpurun_out/gather_2013.json 2> gpurun_out/gather_2013.err; tail -30 gpurun_out/gather_2013.json
timeout 300 python scripts/bench_gather.py --grss2018 > gpurun_out/gather_2018.json 2> gpurun_out/gather_2018.err; tail -30 gpurun_out/gather_2018.json
# 3. ncu of both gather kernels (one launch each): dram bytes, achieved bandwidth, stall reasons
for v in v1 v2; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gather_kernel -c 2 \
    -o gpurun_out/gather_$v -f python scripts/bench_gather.py --reps 1 --only $v > gpurun_out/ncu_gather_$v.log 2>&1
done
# 4. headline bench, both arms
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 python scripts/bench_inference.py > gpurun_out/inference.json 2> gpurun_out/inference.err; tail -1 gpurun_out/inference.json
# 5. whole-scene inference throughput (above)
