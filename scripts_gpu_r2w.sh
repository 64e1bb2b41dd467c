#!/bin/bash
# Round 2, eight GPUs: what the overlapped all-reduce costs the persistent GEMMs -- NCCL CTA budget variants
mkdir -p gpurun_out/r2w
O=gpurun_out/r2w
RUN8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
for v in default 4 2 8; do
  if [ $v = default ]; then unset NCCL_MAX_CTAS; else export NCCL_MAX_CTAS=$v; fi
  timeout 300 $RUN8 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline > $O/bench_8gpu_ctas_$v.log 2>&1; tail -1 $O/bench_8gpu_ctas_$v.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('NCCL_MAX_CTAS=$v', round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done
