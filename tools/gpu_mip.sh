#!/bin/bash
# mip stage: parity first, then the micro-benchmark, then (NCU=1) ncu on the dense builds
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "mip or occupancy or sparse_frame or two_grids or voxelize_bit_exact" ) > $O/pytest_mip.log 2>&1; tail -8 $O/pytest_mip.log
timeout 600 python tools/mip_bench.py ${SIZES:-256 512 1024} > $O/mip_bench.jsonl 2> $O/mip_bench.err; cat $O/mip_bench.jsonl; tail -3 $O/mip_bench.err
if [ -n "$NCU" ]; then
  KINDS=${NCU_KINDS:-scene} timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mip_" -s 6 -c 4 -f -o $O/mip_r2 python tools/mip_bench.py ${NCU_SIZE:-256} > $O/ncu_mip.log 2>&1; tail -2 $O/ncu_mip.log
fi
