#!/bin/bash
# Round 2, GPU call F: fused inference epilogue (parity + throughput vs the two-pass form), whole suite, racecheck
mkdir -p gpurun_out/r2f
O=gpurun_out/r2f
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_dualcnn.py tests/test_gpu_concnn.py tests/test_gpu_end_to_end.py -q -k "eval or infer or predict or end_to_end or rows or dualcnn or concnn" 2>&1 | tail -12 > $O/pytest_eval.log; cat $O/pytest_eval.log | cut -c1-300
timeout 600 python scripts/bench_inference.py > $O/inference_fused.json 2> $O/inference_fused.err; tail -1 $O/inference_fused.json; tail -2 $O/inference_fused.err
HYP_EVAL_UNFUSED=1 timeout 600 python scripts/bench_inference.py > $O/inference_twopass.json 2> $O/inference_twopass.err; tail -1 $O/inference_twopass.json
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/pytest_gpu.log; grep -E "passed|failed|FAILED|ERROR" $O/pytest_gpu.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q -k "cta_pair and (128-64-32 or 256-256-64 or 480-240-1000) or f16x3 and 256-240-480" > $O/sanitizer_racecheck_tc.log 2>&1; tail -4 $O/sanitizer_racecheck_tc.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/one_step.py --steps 1 --batch 256 > $O/sanitizer_memcheck_step.log 2>&1; tail -4 $O/sanitizer_memcheck_step.log
