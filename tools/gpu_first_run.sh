#!/bin/bash
# first GPU contact: parity suite + a quick timing of the benchmark frame
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -40 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python tools/quick_time.py 2>&1 | tee gpurun_out/quick_time.txt
