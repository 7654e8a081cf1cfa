#!/bin/bash
# per-kernel durations of the mip stage with WARM caches (ncu --cache-control none): what the kernels cost inside a frame loop
mkdir -p gpurun_out
O=gpurun_out
for kind in scene random; do
  KINDS=$kind timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"mip_" --csv --log-file $O/mip_warm_$kind.csv python tools/mip_bench.py ${SIZES:-256 1024} > /dev/null 2>&1
  echo "== $kind"; grep -E "mip_" $O/mip_warm_$kind.csv | awk -F'","' '{print substr($5,1,22), $NF}' | tr -d '"' | awk '{n[$1]++; if (n[$1] % 10 == 0) print}' | head -40
done
