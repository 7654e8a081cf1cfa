#!/bin/bash
# ncu --set full of the large-scene front-half kernels (one launch each) at config ${CFG:-4}
mkdir -p gpurun_out
CFG=${CFG:-4}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"vox_setup|cam_setup|vox_resolve|cam_raster|vox_raster|cam_resolve" --launch-skip 12 -c 6 \
  -o gpurun_out/front_c$CFG -f python bench.py --config $CFG --steps 1 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_front_c$CFG.log 2>&1
tail -3 gpurun_out/ncu_front_c$CFG.log | cut -c1-200
ls -la gpurun_out/front_c$CFG.ncu-rep
