#!/bin/bash
# whole GPU suite + mip micro-benchmark + per-kernel times of configs 2, 4, 5
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 600 python tools/mip_bench.py ${SIZES:-256 512 1024} > $O/mip_bench.jsonl 2> $O/mip_bench.err; cat $O/mip_bench.jsonl; tail -3 $O/mip_bench.err
for c in 2 4 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu > $O/launch_c$c.log 2>&1
  python tools/launch_summary.py $O/launches_c$c.csv > $O/launch_summary_c$c.txt 2>&1; cat $O/launch_summary_c$c.txt
done
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_quick.json 2> $O/bench_quick.err; cut -c1-300 $O/bench_quick.json; tail -3 $O/bench_quick.err
