#!/bin/bash
# strong-scaling sweep on one box: N in $NS (default "1 2 4 8"), configs in $CONFIGS (default "2"), exchanges in $EXS (default "p2p nccl")
mkdir -p gpurun_out/scale
nvidia-smi topo -m > gpurun_out/scale/topo.txt 2>&1
for cfg in ${CONFIGS:-2}; do
  for n in ${NS:-1 2 4 8}; do
    for ex in ${EXS:-p2p nccl}; do
      [ $n -eq 1 ] && [ $ex != p2p ] && continue
      out=gpurun_out/scale/c${cfg}_n${n}_$ex
      if [ $n -eq 1 ]; then
        timeout 600 python bench.py --gpus 1 --steps ${STEPS:-50} --warmup 5 --no-cpu --config $cfg > $out.json 2> $out.err
      else
        timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps ${STEPS:-50} --warmup 5 --exchange $ex --config $cfg > $out.json 2> $out.err
      fi
      python - $out.json $cfg $n $ex <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(f"config {sys.argv[2]} N={sys.argv[3]} {sys.argv[4]}: {d['ms_per_step']:.3f} ms/frame  {d['value']:.1f} fps", d.get("stages_us_per_rank", [""])[0])
except Exception as e:
    print(f"config {sys.argv[2]} N={sys.argv[3]} {sys.argv[4]}: FAILED {e}")
PY
      grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$\|NCCL version" $out.err | tail -4
    done
  done
done
