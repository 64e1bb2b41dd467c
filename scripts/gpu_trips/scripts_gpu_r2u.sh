#!/bin/bash
# Round 2, GPU call U: level wgrad with tap groups sharing their activation tiles (TcSeg.nsets)
mkdir -p gpurun_out/r2u
O=gpurun_out/r2u
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_zzz_c5_joint.py -q -x 2>&1 | tail -6 > $O/pytest.log; cat $O/pytest.log | cut -c1-300
for v in 1 2 3 4; do
HYP_WG_TAP_SETS=$v HYP_PROF_LAYERS=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prof-out $O/prof_sets$v.json > $O/prof_sets$v.log 2>&1
python - <<PY
import json
d=json.load(open("$O/prof_sets$v.json"))
print("sets $v", {k.split('/')[-1]: round(x['ms_per_step'],3) for k,x in d.items() if 'wgrad/connector_' in k and 'conv' not in k}, round(sum(x['ms_per_step'] for x in d.values()),3))
PY
done
