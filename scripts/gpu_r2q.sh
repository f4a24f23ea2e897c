#!/bin/bash
# N = 2: sharded tests (incl. pipelined + exact) and bench with / without pipelining
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -15
for mode in "" "--no-pipeline"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 5 --no-knn $mode > gpurun_out/r2q_n2$mode.json 2> gpurun_out/r2q_n2$mode.err; echo "n2 $mode rc=$?"; tail -2 gpurun_out/r2q_n2$mode.err
python -c "
import json
d=json.load(open('gpurun_out/r2q_n2$mode.json'))
print('$mode', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kern_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_vs_n1'], d['clocks']['sm_mhz'])"
done
