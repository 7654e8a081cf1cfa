#!/bin/bash
# per-kernel times (ncu launch list, cold cache, serialised) of configs $CONFIGS (default "4")
mkdir -p gpurun_out
O=gpurun_out
for c in ${CONFIGS:-4}; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $O/launches_c$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu --no-extra --frames-in-flight 1 > $O/launch_c$c.log 2>&1
  python tools/launch_summary.py $O/launches_c$c.csv > $O/launch_summary_c$c.txt 2>&1; cat $O/launch_summary_c$c.txt
done
