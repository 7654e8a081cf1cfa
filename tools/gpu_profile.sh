#!/bin/bash
# ncu full captures of the cone tracer (both samplers) and the mip kernels + sampler accuracy experiment
mkdir -p gpurun_out
timeout 600 python tools/sampler_experiment.py > gpurun_out/sampler.txt 2>&1; cat gpurun_out/sampler.txt
VCT_SAMPLER=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cone_kernel -s 3 -c 1 -o gpurun_out/prof_cone_fp32 -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_cone_fp32.log 2>&1
VCT_SAMPLER=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:cone_kernel -s 3 -c 1 -o gpurun_out/prof_cone_tex -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_cone_tex.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mip_fused|vox_|cam_" -s 27 -c 9 -o gpurun_out/prof_small -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out
