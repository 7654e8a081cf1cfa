#!/bin/bash
# the CUDA path against the reference's own GLSL (oracle/_ref/libvct_glsl_ref.so, shipped prebuilt), a bench line and the reference arm of this tree
mkdir -p gpurun_out
O=gpurun_out
( time timeout 200 python -m pytest tests -m gpu -x -q -k "reference_glsl" ) > $O/pytest_glsl_ref.log 2>&1; tail -4 $O/pytest_glsl_ref.log
timeout 150 python bench.py --steps 100 --warmup 5 --no-extra > $O/bench_glsl.json 2> $O/bench_glsl.err; cut -c1-160 $O/bench_glsl.json; tail -2 $O/bench_glsl.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_glsl.json 2> $O/bench_reference_glsl.err; cut -c1-120 $O/bench_reference_glsl.json
