#!/bin/bash
# parity tests + stage times + launch list
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python tools/quick_time.py > $O/quick_time.txt 2>&1; grep "sampler=1" $O/quick_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/launch_bench.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launch_summary.txt 2>&1; cat $O/launch_summary.txt
