#!/bin/bash
# Round 2, GPU call S: flat bn_apply restored (zig-zag sweeps + ks_last kept), GAN step kernels back to the summation order of the chain
mkdir -p gpurun_out/r2s
O=gpurun_out/r2s
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_full_size.py tests/test_gpu_gan_train.py -q -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-400
timeout 300 python scripts/bench_gan.py --batches 32,256 > $O/gan.json 2> $O/gan.err; tail -5 $O/gan.json | cut -c1-250; tail -2 $O/gan.err
