#!/bin/bash
for q in 0 1; do echo "VCT_QUAD=$q"; VCT_QUAD=$q python tools/cone_variants.py 2>&1 | grep "sampler=1"; done
VCT_QUAD=1 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "frame or cone" 2>&1 | tail -2
