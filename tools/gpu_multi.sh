#!/bin/bash
# multi-GPU: peer-exchange parity test, then the bench at N=1 and N=$1 with both exchanges
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15
run() {  # n exchange config
  if [ $1 -eq 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 100 --warmup 5 --no-cpu --config $3
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $1 --steps 100 --warmup 5 --exchange $2 --config $3
  fi
}
for cfg in ${CONFIGS:-2}; do
  run 1 p2p $cfg > gpurun_out/bench_c${cfg}_n1.json 2> gpurun_out/bench_c${cfg}_n1.err; echo "== config $cfg N=1"; cut -c1-160 gpurun_out/bench_c${cfg}_n1.json; tail -3 gpurun_out/bench_c${cfg}_n1.err
  for n in ${NS:-$N}; do
    for ex in p2p nccl; do
      run $n $ex $cfg > gpurun_out/bench_c${cfg}_n${n}_$ex.json 2> gpurun_out/bench_c${cfg}_n${n}_$ex.err
      echo "== config $cfg N=$n $ex"; cut -c1-160 gpurun_out/bench_c${cfg}_n${n}_$ex.json; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_c${cfg}_n${n}_$ex.err | tail -5
    done
  done
done
