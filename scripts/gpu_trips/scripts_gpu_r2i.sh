#!/bin/bash
# Round 2, GPU call I: elementwise kernel changes (parity + bench), whole suite
mkdir -p gpurun_out/r2i
O=gpurun_out/r2i
timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -12 > $O/pytest_gpu.log; cat $O/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
bash scripts/ncu_step.sh; cp gpurun_out/ncu/launches.csv $O/launches_3xf16.csv
