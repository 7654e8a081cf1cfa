#!/bin/bash
# time cone-kernel variants (VARIANTS="2 6 ...") on the three Cornell configs
mkdir -p gpurun_out
for v in ${VARIANTS:-2}; do VCT_CONE_VARIANT=$v timeout 300 python tools/cone_variants.py > gpurun_out/variant_$v.txt 2>&1; grep "sampler=1" gpurun_out/variant_$v.txt; done
