#!/bin/bash
# parity tests, then time cone-kernel variants (VARIANTS="2 3 ...") on the three Cornell configs
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-2}; do VCT_CONE_VARIANT=$v timeout 300 python tools/cone_variants.py > gpurun_out/variant_$v.txt 2>&1; grep "sampler=1" gpurun_out/variant_$v.txt; done
