#!/bin/bash
# one GPU box visit for the tracked evidence: parity tests, smoke, bench (ours + reference arm), stage times, launch list, ncu --set full captures
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py --steps 200 --warmup 10 > $O/bench.json 2> $O/bench.err; cut -c1-300 $O/bench.json; tail -3 $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; cut -c1-300 $O/bench_reference.json
timeout 300 python tools/quick_time.py > $O/quick_time.txt 2>&1; cat $O/quick_time.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/launch_bench.log 2>&1
python tools/launch_summary.py $O/launches.csv > $O/launch_summary.txt 2>&1; cat $O/launch_summary.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:cone_kernel -s 3 -c 1 -f -o $O/cone_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_cone.log 2>&1
timeout 600 $NCU -k regex:"mip_" -s 6 -c 2 -f -o $O/mip_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_mip.log 2>&1
VCT_MIP_DENSE=1 timeout 600 $NCU -k regex:"mip_" -s 6 -c 2 -f -o $O/mip_dense_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_mip_dense.log 2>&1
timeout 600 $NCU -k regex:"vox_|cam_|sparse_|fill_u64|tile_list|shade" -s 33 -c 11 -f -o $O/small_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_small.log 2>&1
ls -la $O | head -40
