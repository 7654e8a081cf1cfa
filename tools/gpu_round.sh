#!/bin/bash
# GPU session helper: parity suite, per-stage timings, bench (both arms), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/pytest_gpu.txt
cat gpurun_out/pytest_gpu.txt
timeout 300 python tools/quick_time.py > gpurun_out/quick_time.txt 2>&1; cat gpurun_out/quick_time.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
