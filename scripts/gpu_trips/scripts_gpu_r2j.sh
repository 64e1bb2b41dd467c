#!/bin/bash
# Round 2, GPU call J: fused CycleGAN step kernels (tests + bench), shifted-load bn-backward apply
mkdir -p gpurun_out/r2j
O=gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_gan_train.py tests/test_gpu_gan.py tests/test_gpu_zz_gan_session.py -q -x 2>&1 | tail -15 > $O/pytest_gan.log; grep -E "passed|failed|Error|error|assert" $O/pytest_gan.log | head -20 | cut -c1-300
timeout 300 python scripts/bench_gan.py > $O/gan_fused.json 2> $O/gan_fused.err; cat $O/gan_fused.json | cut -c1-330; tail -3 $O/gan_fused.err
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/bench_gan.py --batches 32 --steps 2 --warmup 1 > $O/sanitizer_memcheck_gan.log 2>&1; tail -3 $O/sanitizer_memcheck_gan.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -k "3xf16 or 3xtf32" 2>&1 | tail -4 > $O/pytest_parity.log; cat $O/pytest_parity.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
