#!/bin/bash
# ncu --set full on selected GEMM launches of the 2nd train step; exports CSV pages and drops the big .ncu-rep
# usage: scripts/ncu_gemm.sh <name>:<gemm launch index in step> ...
mkdir -p gpurun_out/ncu
for spec in "$@"; do
  name=${spec%%:*}; idx=${spec##*:}
  skip=$((59 + idx))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s $skip -c 1 \
      -o gpurun_out/ncu/$name -f python scripts/one_step.py --steps 2 > gpurun_out/ncu/$name.log 2>&1
  ncu -i gpurun_out/ncu/$name.ncu-rep --page raw --csv > gpurun_out/ncu/$name.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page source --csv > gpurun_out/ncu/$name.source.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$name.ncu-rep --page details > gpurun_out/ncu/$name.details.txt 2>/dev/null
  rm -f gpurun_out/ncu/$name.ncu-rep
done
du -sh gpurun_out
