#!/bin/bash
# large-scene paths: parity (synthetic miniature + full-size configs 4, 5), then per-kernel times of configs 4 and 5
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "synthetic or config4 or config5 or gbuffer or voxelize" ) > $O/pytest_big.log 2>&1; tail -6 $O/pytest_big.log
for c in ${CONFIGS:-4 5}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu --no-extra > $O/launch_c$c.log 2>&1
  python tools/launch_summary.py $O/launches_c$c.csv > $O/launch_summary_c$c.txt 2>&1; cat $O/launch_summary_c$c.txt
done
