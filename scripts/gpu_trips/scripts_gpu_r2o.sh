#!/bin/bash
# Round 2, GPU call O: bn_apply without the 64-bit division, fused GAN discriminator step with block-parallel outer products
mkdir -p gpurun_out/r2o
O=gpurun_out/r2o
timeout 900 python -m pytest tests/test_gpu_gan_train.py tests/test_gpu_parity.py tests/test_gpu_dualcnn.py tests/test_gpu_concnn.py -q -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
timeout 300 python scripts/bench_gan.py --batches 32,256,1024 > $O/gan.json 2> $O/gan.err; cut -c1-200 $O/gan.json; tail -2 $O/gan.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precision bf16 > $O/bench_bf16.log 2>&1; tail -1 $O/bench_bf16.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('bf16', d['ms_per_step'], d['roofline']['frac'], d['kernel_breakdown_ms_per_step'])"
