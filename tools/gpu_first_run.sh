#!/bin/bash
# parity suite + quick timings (development helper)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
VCT_SAMPLER=1 timeout 1500 python -m pytest tests -x -q -m gpu -k "frame or tile or determin" 2>&1 | tail -15
timeout 300 python tools/sampler_experiment.py 2>&1 | tee gpurun_out/sampler.txt
