#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --no-header --tb=line 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.log
