#!/bin/bash
# bench + ncu launch list + full capture of the cone tracer (development helper)
mkdir -p gpurun_out
export VCT_SAMPLER=${VCT_SAMPLER:-1}
python bench.py --steps 50 --warmup 3 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['stages'], d['roofline']['gsamples_per_s'], d['clocks'])
PY
tail -5 gpurun_out/bench.err
ncu --set full --clock-control none --import-source on -k regex:cone_trace -s 3 -c 1 -o gpurun_out/prof_trace -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_trace.log 2>&1
