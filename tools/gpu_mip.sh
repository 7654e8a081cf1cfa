#!/bin/bash
# mip stage: parity first, then the micro-benchmark, then ncu on the dense builds, then the whole GPU suite
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mip or occupancy or sparse_frame or two_grids" ) > $O/pytest_mip.log 2>&1; tail -15 $O/pytest_mip.log
timeout 300 python tools/mip_bench.py ${SIZES:-256 512} > $O/mip_bench.jsonl 2> $O/mip_bench.err; cat $O/mip_bench.jsonl; tail -3 $O/mip_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"mip_" -s 8 -c 4 -f -o $O/mip_r2 python tools/mip_bench.py 256 > $O/ncu_mip.log 2>&1; tail -2 $O/ncu_mip.log
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -6 $O/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > $O/bench_quick.json 2> $O/bench_quick.err; cut -c1-400 $O/bench_quick.json; tail -3 $O/bench_quick.err
