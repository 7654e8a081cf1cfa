#!/bin/bash
# mip stage after the rework + per-kernel times of the large configs (ncu launch list: serialised, cold cache)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mip or occupancy or sparse_frame or two_grids" ) > $O/pytest_mip.log 2>&1; tail -5 $O/pytest_mip.log
timeout 600 python tools/mip_bench.py ${SIZES:-256 512 1024} > $O/mip_bench.jsonl 2> $O/mip_bench.err; cat $O/mip_bench.jsonl; tail -3 $O/mip_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mip_" -s 8 -c 6 -f -o $O/mip_r2 python tools/mip_bench.py 256 > $O/ncu_mip.log 2>&1; tail -2 $O/ncu_mip.log
for c in 4 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu > $O/launch_c$c.log 2>&1
  python tools/launch_summary.py $O/launches_c$c.csv > $O/launch_summary_c$c.txt 2>&1; cat $O/launch_summary_c$c.txt
done
