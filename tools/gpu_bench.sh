#!/bin/bash
# the driver's two bench commands (timed), the cone-kernel ncu capture that bench.py's roofline reads, smoke
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:cone_kernel -s 3 -c 1 -f -o $O/cone_full python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/ncu_cone.log 2>&1; tail -1 $O/ncu_cone.log
python tools/ncu_to_json.py $O/cone_full.ncu-rep config2_sampler1 $O/r02_cone_kernel_ncu.json > /dev/null && cp $O/r02_cone_kernel_ncu.json profiles/r02_cone_kernel_ncu.json
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real; cut -c1-200 $O/bench.json; tail -3 $O/bench.err
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/bench_reference.json 2> $O/bench_reference.err ) 2>&1 | grep real; cut -c1-300 $O/bench_reference.json; tail -3 $O/bench_reference.err
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
