for mb in 48 96 128 256; do
  HYP_WG_SLICE_MB=$mb HYP_PROF_LAYERS=1 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --prof-out gpurun_out/prof_$mb.json > gpurun_out/bench_$mb.log 2>&1
  python - <<PY
import json
d=json.load(open('gpurun_out/prof_$mb.json'))
b=json.loads(open('gpurun_out/bench_$mb.log').read().strip().split('\n')[-1])
print($mb, 'ms/step', round(b['ms_per_step'],3), {k.split('/')[-1]:round(v['ms_per_step'],3) for k,v in d.items() if 'wgrad/connector_' in k and 'conv' not in k})
PY
done
