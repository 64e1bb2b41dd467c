#!/bin/bash
# Round 2, GPU call U2: bn_apply sweeps its rows descending (starts on the z rows the GEMM wrote last)
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -3 > $O/pytest_rev.log; cat $O/pytest_rev.log | cut -c1-200
for z in 1 0 1 0; do
HYP_SWEEP_ZIGZAG=$z timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_zz$z.log 2>&1; tail -1 $O/bench_zz$z.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('zigzag $z', round(d['ms_per_step'],3), d['kernel_breakdown_ms_per_step']['tc_bn_apply_kernel'], d['kernel_breakdown_ms_per_step']['tc_gemm_kernel'], d['clocks']['sm_mhz'])"
done
