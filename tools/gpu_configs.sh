#!/bin/bash
# single-GPU bench of the larger BASELINE configs (CONFIGS="3 4 5"), fewer steps
mkdir -p gpurun_out
for cfg in ${CONFIGS:-3 4 5}; do
  timeout 900 python bench.py --gpus 1 --steps ${STEPS:-30} --warmup 3 --no-cpu --config $cfg > gpurun_out/bench_c${cfg}_n1.json 2> gpurun_out/bench_c${cfg}_n1.err
  echo "== config $cfg"; python - <<PY
import json
d=json.load(open("gpurun_out/bench_c${cfg}_n1.json"))
print(d["ms_per_step"], d["config"]["workload"][:60], {k: round(v,1) for k,v in d["stages"].items() if k.endswith("_us")}, d["roofline"]["gsamples_per_s"])
PY
  tail -2 gpurun_out/bench_c${cfg}_n1.err
done
