#!/bin/bash
# parity tests + stage times + ncu --set full of the small kernels (voxelizer, G-buffer, mip)
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $O/pytest_gpu.log 2>&1; tail -4 $O/pytest_gpu.log
timeout 300 python tools/quick_time.py > $O/quick_time.txt 2>&1; grep "sampler=1" $O/quick_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"vox_|cam_|mip_|scan_|fill_|tile_list|shade" -s 36 -c 13 -f -o $O/small_full python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_small.log 2>&1
tail -3 $O/ncu_small.log
