#!/bin/bash
# frames in flight: bench at N in $NS (default "1 2"), F in $FS (default "1 2"), config 2, headline only
mkdir -p gpurun_out/fif
for n in ${NS:-1 2}; do
  for f in ${FS:-1 2}; do
    out=gpurun_out/fif/c2_n${n}_f$f
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps ${STEPS:-100} --warmup 5 --no-cpu --no-extra --frames-in-flight $f $EXTRA > $out.json 2> $out.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-100} --warmup 5 --no-extra --frames-in-flight $f $EXTRA > $out.json 2> $out.err
    fi
    python - $out.json $n $f <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"N={sys.argv[2]} F={sys.argv[3]}: {d['ms_per_step']:.3f} ms/frame  {d['value']:.1f} fps  e2e {d['e2e']['ms_per_step']:.3f} ms", d.get("stages_us_per_rank", [""])[0])
except Exception as e:
    print(f"N={sys.argv[2]} F={sys.argv[3]}: FAILED {e}")
PY
    grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" $out.err | tail -4
  done
done
