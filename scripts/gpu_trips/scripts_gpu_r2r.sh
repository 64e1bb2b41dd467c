#!/bin/bash
# Round 2, GPU call R: 2-D bn_apply, zig-zag sweep directions, partial last K blocks (ks_last); fused GAN gs_dot change
mkdir -p gpurun_out/r2r
O=gpurun_out/r2r
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_full_size.py tests/test_gpu_gan_train.py -q -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-700
HYP_SWEEP_ZIGZAG=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_nozz.log 2>&1; tail -1 $O/bench_nozz.log | cut -c1-300
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers.json > $O/prof.log 2>&1; tail -1 $O/prof.log | cut -c1-200
HYP_SWEEP_ZIGZAG=0 HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers_nozz.json > $O/prof_nozz.log 2>&1; tail -1 $O/prof_nozz.log | cut -c1-200
timeout 300 python scripts/bench_gan.py --batches 32,256,1024 > $O/gan.json 2> $O/gan.err; tail -5 $O/gan.json | cut -c1-400; tail -2 $O/gan.err
