#!/bin/bash
# parity tests, then time cone-kernel variants on the three Cornell configs: VARIANTS="2 3" x PERSIST="1 0"
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for pe in ${PERSIST:-1}; do for v in ${VARIANTS:-3}; do echo "VCT_CONE_PERSIST=$pe VCT_CONE_VARIANT=$v"; VCT_CONE_PERSIST=$pe VCT_CONE_VARIANT=$v timeout 300 python tools/cone_variants.py 2>&1 | grep "sampler=1"; done; done
