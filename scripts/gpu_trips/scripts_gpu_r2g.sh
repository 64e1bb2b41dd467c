#!/bin/bash
# Round 2, GPU call G: ncu --set full of three GEMM launches in 3xF16 mode (1x1 forward, 1x1 dgrad, level forward),
# C3 bench, headline bench
mkdir -p gpurun_out/r2g gpurun_out/ncu
O=gpurun_out/r2g
bash scripts/ncu_gemm.sh fwd_conv_enc_2:2 dgrad_1x1:53 fwd_connector_1:8 > $O/ncu_gemm.log 2>&1; tail -2 $O/ncu_gemm.log
cp gpurun_out/ncu/fwd_conv_enc_2.* gpurun_out/ncu/dgrad_1x1.* gpurun_out/ncu/fwd_connector_1.* $O/ 2>/dev/null
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-600
timeout 600 python bench.py --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51.log 2>&1; tail -1 $O/bench_c3_51.log | cut -c1-600
timeout 600 python bench.py --workload c3_grss2018 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_49.log 2>&1; tail -1 $O/bench_c3_49.log | cut -c1-600
