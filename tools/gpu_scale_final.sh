#!/bin/bash
# the driver's scaling command at N = $1 (both extra configs included), line kept under gpurun_out/scale/
N=$1
mkdir -p gpurun_out/scale
out=gpurun_out/scale/final_n$N
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-50} --warmup 5 > $out.json 2> $out.err ) 2>&1 | grep real
python - $out.json <<'PY'
import json, sys
d = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")][0]
print(f"N={d['n_gpus']}: {d['ms_per_step']:.4f} ms/frame {d['value']:.1f} fps, e2e {d['e2e']['ms_per_step']:.4f} ms, F={d['config']['frames_in_flight']}")
for k, v in d.get("extra_configs", {}).items():
    print("  config", k, v.get("ms_per_frame"), v.get("error"))
print("  rank 0 stages", d.get("stages_us_per_rank", [None])[0])
PY
grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" $out.err | tail -3
