#!/bin/bash
# Round 2, GPU call K: after the bn-backward kernel fixes — parity, bench in the three precisions, per-layer profile,
# launch list, role timing, GAN bench, inference, C3
mkdir -p gpurun_out/r2k gpurun_out/ncu
O=gpurun_out/r2k
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest_gpu.log; grep -E "passed|failed|FAILED|ERROR" $O/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --precision 3xtf32 > $O/bench_tf32.log 2>&1; tail -1 $O/bench_tf32.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 > $O/bench_bf16.log 2>&1; tail -1 $O/bench_bf16.log | cut -c1-200
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers.json > $O/bench_layers.log 2>&1
timeout 300 python scripts/bench_gan.py --batches 32,256,1024,16384 > $O/gan_graphs.json 2> $O/gan.err; cut -c1-200 $O/gan_graphs.json
timeout 600 python scripts/bench_inference.py > $O/inference.json 2> $O/inference.err; tail -1 $O/inference.json
timeout 600 python bench.py --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51.log 2>&1; tail -1 $O/bench_c3_51.log | cut -c1-300
timeout 600 python bench.py --workload c3_grss2018 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_49.log 2>&1; tail -1 $O/bench_c3_49.log | cut -c1-300
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/tc_timing.log 2>&1
bash scripts/ncu_step.sh; cp gpurun_out/ncu/launches.csv $O/launches_3xf16.csv
bash scripts/ncu_gemm.sh fwd_conv_enc_2:2 dgrad_1x1:53 fwd_connector_1:8 > $O/ncu_gemm.log 2>&1; cp gpurun_out/ncu/fwd_conv_enc_2.raw.csv gpurun_out/ncu/dgrad_1x1.raw.csv gpurun_out/ncu/fwd_connector_1.raw.csv $O/ 2>/dev/null
