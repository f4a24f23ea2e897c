#!/bin/bash
# N = 2: sharded parity tests + short bench at N=1 and N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-knn --no-data-sweep > gpurun_out/r2i_n1.json 2> gpurun_out/r2i_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-knn > gpurun_out/r2i_n2.json 2> gpurun_out/r2i_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r2i_n2.err
python -c "
import json
for f in ('r2i_n1','r2i_n2'):
    d=json.load(open(f'gpurun_out/{f}.json'))
    print(f, d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kern_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_vs_n1'], d['clocks']['sm_mhz'])
"
