#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/quick_time.py > gpurun_out/quick_time.txt 2>&1; grep "sampler=1" gpurun_out/quick_time.txt
