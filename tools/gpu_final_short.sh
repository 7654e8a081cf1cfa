#!/bin/bash
# refresh of the tracked evidence after a cone-kernel change: bench, launch list, ncu --set full of the cone kernel
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 200 --warmup 10 > $O/bench.json 2> $O/bench.err; cut -c1-200 $O/bench.json; tail -3 $O/bench.err
timeout 300 python tools/quick_time.py > $O/quick_time.txt 2>&1; grep "sampler=1" $O/quick_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/launch_bench.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launch_summary.txt 2>&1; cat $O/launch_summary.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cone_kernel -s 3 -c 1 -f -o $O/cone_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_cone.log 2>&1
