#!/bin/bash
# Round 2, eight GPUs: SMs left to the collective (HYP_TC_SMS) x NCCL CTA budget
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
for v in 144:4 140:8 132:default; do
  sms=${v%%:*}; ctas=${v##*:}
  if [ $ctas = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$ctas; fi
  HYP_TC_SMS=$sms timeout 300 $RUN8 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_8gpu_sms_${sms}_ctas_$ctas.log 2>&1; tail -1 $O/bench_8gpu_sms_${sms}_ctas_$ctas.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('HYP_TC_SMS=$sms NCCL_MAX_CTAS=$ctas', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done
