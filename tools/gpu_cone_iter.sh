#!/bin/bash
# quick iteration: all GPU parity tests, cone-kernel variant timings, ncu --set full of the cone kernel
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
for v in ${VARIANTS:-2 3 4}; do VCT_CONE_VARIANT=$v timeout 300 python tools/cone_variants.py > $O/variant_$v.txt 2>&1; grep "sampler=1" $O/variant_$v.txt; done
timeout 300 python tools/quick_time.py > $O/quick_time.txt 2>&1; grep "sampler=1" $O/quick_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cone_kernel -s 3 -c 1 -f -o $O/cone_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_cone.log 2>&1
ls -la $O | head -30
