#!/bin/bash
# warp-state / scheduler breakdown of the cone kernel (why 0.77 of the TEX roofline): one launch, sections only, raw CSV
mkdir -p gpurun_out
O=gpurun_out
timeout 240 ncu --section WarpStateStats --section SchedulerStats --section Occupancy --section LaunchStats --section InstructionStats \
  --metrics l1tex__texin_sm2tex_req_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_tex_wavefronts.sum,l1tex__f_wavefronts.sum,sm__inst_executed_pipe_tex.sum,smsp__inst_executed_pipe_tex.sum,l1tex__t_requests_pipe_tex_mem_texture.sum,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max,smsp__thread_inst_executed_per_inst_executed.ratio \
  --clock-control none -k regex:cone_kernel -s 3 -c 1 --csv --page raw --log-file $O/cone_stalls.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-extra > $O/cone_stalls.log 2>&1
tail -2 $O/cone_stalls.log | cut -c1-200; wc -c $O/cone_stalls.csv
