#!/bin/bash
# Round 2, GPU call E: GAN session tests + graph replay, GAN bench (graphs vs eager), dgrad reductions vs RMW,
# role timing of the GEMM launches in 3xF16 mode
mkdir -p gpurun_out/r2e
O=gpurun_out/r2e
timeout 900 python -m pytest tests/test_gpu_zz_gan_session.py tests/test_gpu_gan_train.py tests/test_gpu_zzz_c5_joint.py -q 2>&1 | tail -30 > $O/pytest_gan.log; cat $O/pytest_gan.log | cut -c1-300
timeout 300 python scripts/bench_gan.py > $O/gan_graphs.json 2> $O/gan_graphs.err; cat $O/gan_graphs.json | cut -c1-400; tail -3 $O/gan_graphs.err
HYP_GAN_GRAPHS=0 timeout 300 python scripts/bench_gan.py > $O/gan_eager.json 2> $O/gan_eager.err; cat $O/gan_eager.json | cut -c1-400
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_red.log 2>&1; tail -1 $O/bench_red.log | cut -c1-400
HYP_DGRAD_RMW=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_rmw.log 2>&1; tail -1 $O/bench_rmw.log | cut -c1-400
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/tc_timing.log 2>&1; grep -c tc_timing $O/tc_timing.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -k "3xf16 or full or large or doubled or eval_rows or bf16" 2>&1 | tail -8 > $O/pytest_parity.log; cat $O/pytest_parity.log | cut -c1-300
