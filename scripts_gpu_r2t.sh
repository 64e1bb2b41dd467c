#!/bin/bash
# Round 2, GPU call T: epilogue slabs through the TMA unit (tile store / f32 add reduction)
mkdir -p gpurun_out/r2t
O=gpurun_out/r2t
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_parity.py tests/test_gpu_full_size.py -q -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench.log 2>&1; tail -1 $O/bench.log | cut -c1-330
HYP_TC_TMA_STORE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_plain.log 2>&1; tail -1 $O/bench_plain.log | cut -c1-330
for d in 1 0; do
HYP_TC_TMA_STORE=$d HYP_TC_TIMING=1 timeout 300 python scripts/one_step.py --steps 2 > $O/timing_tma$d.log 2>&1
grep "tc_timing" $O/timing_tma$d.log | tail -59 | grep -E "fwd/conv_enc_2|fwd/conv_dec_0|fwd/conv_enc_1|dgrad/conv_dec_0|dgrad/conv_enc_2|dgrad/connector_1|fwd/fc_0" | sed "s/^/tma$d /" | cut -c1-330
done
