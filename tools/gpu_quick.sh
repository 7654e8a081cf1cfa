#!/bin/bash
# parity suite + per-stage timings (development helper)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python tools/quick_time.py > gpurun_out/quick_time.txt 2>&1; cat gpurun_out/quick_time.txt
