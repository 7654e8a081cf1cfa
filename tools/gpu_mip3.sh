#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mip or occupancy or sparse_frame or two_grids or voxelize" ) > $O/pytest_mip.log 2>&1; tail -5 $O/pytest_mip.log
timeout 600 python tools/mip_bench.py ${SIZES:-256 512 1024} > $O/mip_bench.jsonl 2> $O/mip_bench.err; cat $O/mip_bench.jsonl; tail -3 $O/mip_bench.err
KINDS=scene timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mip_" -s 6 -c 4 -f -o $O/mip_scene python tools/mip_bench.py 256 > $O/ncu_mip.log 2>&1; tail -2 $O/ncu_mip.log
KINDS=scene timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_mip.csv python tools/mip_bench.py 256 512 > /dev/null 2>&1; grep -E "mip_" $O/launches_mip.csv | awk -F'","' '{print substr($5,1,20), $NF}' | sort | uniq -c | sort -k2,2 -k3,3n | awk '{print}' | head -60 | cut -c1-60 | tr '\n' ';'
