#!/bin/bash
# Round 2, GPU call C: the 16-bit operand formats — GEMM building block first, then the engine parity matrix, then bench
mkdir -p gpurun_out/r2c
O=gpurun_out/r2c
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -15 > $O/pytest_tc.log; cat $O/pytest_tc.log
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -k "not fp32" 2>&1 | tail -40 > $O/pytest_parity.log; cat $O/pytest_parity.log
for prec in 3xtf32 3xf16 bf16; do
  timeout 600 python bench.py --steps 10 --warmup 3 --precision $prec > $O/bench_$prec.log 2>&1; tail -1 $O/bench_$prec.log | cut -c1-1200
done
