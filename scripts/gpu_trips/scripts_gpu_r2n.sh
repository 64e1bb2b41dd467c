#!/bin/bash
# Round 2, GPU call N: 3xF16 without value planes on the internal tensors — parity + bench
mkdir -p gpurun_out/r2n
O=gpurun_out/r2n
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_end_to_end.py tests/test_gpu_zz_classify_apps.py tests/test_gpu_zzz_c5_joint.py -q -x 2>&1 | tail -8 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['config']['l2'][:40], d['parity_vs_fp64_oracle'], d['kernel_breakdown_ms_per_step'])"
