#!/bin/bash
# Round 2, final single-GPU evidence trip: whole device suite + smoke, bench lines (three precisions, C3, gather),
# per-kernel profile, role timing, ncu launch list + --set full captures, GAN / inference benches, compute-sanitizer.
# scripts/make_profiles_r02.py turns gpurun_out/r2z into profiles/r02_*.
mkdir -p gpurun_out/r2z gpurun_out/ncu
O=gpurun_out/r2z
timeout 2400 python -m pytest tests -m gpu -q 2>&1 | tail -12 > $O/pytest_gpu.log; grep -E "passed|failed|FAILED|ERROR" $O/pytest_gpu.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench.log 2>&1; tail -1 $O/bench.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['kernel_breakdown_ms_per_step'])"
timeout 300 python bench.py --steps 10 --warmup 3 --precision 3xtf32 > $O/bench_tf32.log 2>&1; tail -1 $O/bench_tf32.log | cut -c1-200
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16 > $O/bench_bf16.log 2>&1; tail -1 $O/bench_bf16.log | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.log 2>&1; tail -1 $O/bench_reference.log | cut -c1-300
HYP_PROF_LAYERS=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_layers.json > $O/bench_layers.log 2>&1
timeout 300 python scripts/bench_gan.py --batches 32,256,1024,16384 > $O/gan.json 2> $O/gan.err; cut -c1-200 $O/gan.json
HYP_GAN_FUSED=0 timeout 300 python scripts/bench_gan.py --batches 32,256,1024 > $O/gan_chain.json 2> $O/gan_chain.err; cut -c1-200 $O/gan_chain.json
timeout 600 python scripts/bench_inference.py > $O/inference.json 2> $O/inference.err; tail -1 $O/inference.json | cut -c1-300
timeout 600 python bench.py --workload c3_grss2018_51 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_51.log 2>&1; tail -1 $O/bench_c3_51.log | cut -c1-300
timeout 600 python bench.py --workload c3_grss2018 --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_c3_49.log 2>&1; tail -1 $O/bench_c3_49.log | cut -c1-300
timeout 300 python bench.py --workload gather_c2 --steps 50 --warmup 5 > $O/bench_gather_c2.log 2>&1; tail -1 $O/bench_gather_c2.log | cut -c1-300
timeout 300 python bench.py --workload gather_c2 --batch 65536 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_gather_c2_64k.log 2>&1; tail -1 $O/bench_gather_c2_64k.log | cut -c1-300
timeout 300 python bench.py --workload gather_c3 --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_gather_c3.log 2>&1; tail -1 $O/bench_gather_c3.log | cut -c1-300
timeout 300 python scripts/bench_gather.py > $O/gather_2013.json 2> $O/gather_2013.err; tail -2 $O/gather_2013.err
timeout 300 python scripts/bench_gather.py --grss2018 > $O/gather_2018.json 2> $O/gather_2018.err; tail -2 $O/gather_2018.err
timeout 120 python scripts/bench_hbm_mix.py > $O/hbm_mix.json 2>&1; cat $O/hbm_mix.json
HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/tc_timing.log 2>&1
bash scripts/ncu_step.sh; cp gpurun_out/ncu/launches.csv $O/launches_3xf16.csv
bash scripts/ncu_gemm.sh fwd_conv_enc_2:2 dgrad_1x1:53 fwd_connector_1:8 > $O/ncu_gemm.log 2>&1; cp gpurun_out/ncu/fwd_conv_enc_2.raw.csv gpurun_out/ncu/dgrad_1x1.raw.csv gpurun_out/ncu/fwd_connector_1.raw.csv $O/ 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:gather_rows_kernel -c 1 -o $O/gather -f python bench.py --workload gather_c2 --steps 2 --warmup 1 --no-cpu-baseline > $O/ncu_gather.log 2>&1
ncu -i $O/gather.ncu-rep --page raw --csv > $O/gather_raw.csv 2>/dev/null; rm -f $O/gather.ncu-rep
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q > $O/sanitizer_memcheck_tc.log 2>&1; tail -3 $O/sanitizer_memcheck_tc.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/one_step.py --steps 1 --batch 256 > $O/sanitizer_memcheck_step.log 2>&1; tail -3 $O/sanitizer_memcheck_step.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -x -q -k "cta_pair and (128-64-32 or 256-256-64 or 480-240-1000) or f16x3 and 256-240-480" > $O/sanitizer_racecheck_tc.log 2>&1; tail -3 $O/sanitizer_racecheck_tc.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/one_step.py --steps 1 --batch 128 > $O/sanitizer_racecheck_step.log 2>&1; tail -3 $O/sanitizer_racecheck_step.log
du -sh gpurun_out
