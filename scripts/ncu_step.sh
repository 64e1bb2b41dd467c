#!/bin/bash
# ncu launch list of full train steps (device time + DRAM bytes per launch; cold-cache, serialised: compare shares)
mkdir -p gpurun_out/ncu
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv \
    --log-file gpurun_out/ncu/launches.csv python scripts/one_step.py --steps 3 > gpurun_out/ncu/launches.log 2>&1
tail -2 gpurun_out/ncu/launches.log
wc -l gpurun_out/ncu/launches.csv
