#!/bin/bash
# multi-GPU round: sharded parity test + bench at N = 1 and N = all
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "n1 rc=$?"; cut -c1-400 gpurun_out/bench_n1.json
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "n$n rc=$?"; tail -3 gpurun_out/bench_n$n.err; cat gpurun_out/bench_n$n.json | cut -c1-3000
  fi
done
if [ "$N" = "8" ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 scripts/bench_config5.py > gpurun_out/config5_n8.json 2> gpurun_out/config5_n8.err; echo "config5 rc=$?"; tail -2 gpurun_out/config5_n8.err; cat gpurun_out/config5_n8.json
fi
