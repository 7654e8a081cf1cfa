#!/bin/bash
# one GPU: the cone-kernel ncu capture bench.py's roofline reads, both bench arms as the driver runs them, smoke, the launch list of config 2
mkdir -p gpurun_out
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:cone_kernel -s 3 -c 1 -f -o $O/cone_full python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --frames-in-flight 1 > $O/ncu_cone.log 2>&1; tail -1 $O/ncu_cone.log | cut -c1-100
python tools/ncu_to_json.py $O/cone_full.ncu-rep config2_sampler1 $O/r02_cone_kernel_ncu.json > /dev/null && cp $O/r02_cone_kernel_ncu.json profiles/r02_cone_kernel_ncu.json
( time timeout 900 python bench.py > $O/bench.json 2> $O/bench.err ) 2>&1 | grep real; cut -c1-200 $O/bench.json; tail -3 $O/bench.err
( time timeout 900 python bench.py --frames-in-flight 1 --no-extra --no-cpu > $O/bench_f1.json 2> $O/bench_f1.err ) 2>&1 | grep real; cut -c1-200 $O/bench_f1.json
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err ) 2>&1 | grep real; cut -c1-200 $O/bench_reference.json; tail -3 $O/bench_reference.err
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file $O/launches_c2.csv python bench.py --steps 5 --warmup 3 --no-cpu --no-extra --frames-in-flight 1 > $O/launch_c2.log 2>&1
python tools/launch_summary.py $O/launches_c2.csv > $O/launch_summary_c2.txt 2>&1; cat $O/launch_summary_c2.txt
