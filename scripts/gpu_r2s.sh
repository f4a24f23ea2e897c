#!/bin/bash
# N = 8: pipelined vs plain exchange, scan only
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
for mode in "" "--no-pipeline"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 200 --warmup 5 --no-knn --no-config5 $mode > gpurun_out/r2s_n8$mode.json 2> gpurun_out/r2s_n8$mode.err; echo "n8 $mode rc=$?"; tail -2 gpurun_out/r2s_n8$mode.err
python -c "
import json
d=json.load(open('gpurun_out/r2s_n8$mode.json'))
print('$mode', d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kern_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_vs_n1'], d['clocks']['sm_mhz'])"
done
