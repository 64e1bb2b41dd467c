#!/bin/bash
# Round 2, GPU call P: pattern-specialised residual loads in bn_apply, weight scale 2^6; smoke; whole suite
mkdir -p gpurun_out/r2p
O=gpurun_out/r2p
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 > $O/pytest_gpu.log; grep -E "passed|failed|FAILED|ERROR" $O/pytest_gpu.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['parity_vs_fp64_oracle']['max_abs_logit_error'], d['kernel_breakdown_ms_per_step'])"
