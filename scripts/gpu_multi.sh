#!/bin/bash
# N = 8 box: sharded tests, bench at N = 1, 2, 4, 8 (the last with config 5 and the kNN build)
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-knn --no-data-sweep > gpurun_out/r2n_n1.json 2> gpurun_out/r2n_n1.err; echo "n1 rc=$?"
for n in 2 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 100 --warmup 5 --no-knn > gpurun_out/r2n_n$n.json 2> gpurun_out/r2n_n$n.err; echo "n$n rc=$?"; tail -2 gpurun_out/r2n_n$n.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 100 --warmup 5 > gpurun_out/r2n_n8.json 2> gpurun_out/r2n_n8.err; echo "n8 rc=$?"; tail -3 gpurun_out/r2n_n8.err
python -c "
import json
for f in ('r2n_n1','r2n_n2','r2n_n4','r2n_n8'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
    except Exception as e:
        print(f, 'no json', e); continue
    print(f, d['n_gpus'], 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'kern_ms', round(d['roofline']['kernel_ms_avg'],4), 'frac', round(d['roofline']['frac'],3), 'parity', d['parity_vs_n1'], d['clocks']['sm_mhz'])
    if 'config5_100Mx768' in d: print(json.dumps(d['config5_100Mx768']))
    if 'knn_build' in d: print(json.dumps(d['knn_build']))
"
