#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_scan_gpu.py tests/test_exact_gpu.py tests/test_fullsize_gpu.py tests/test_indices_gpu.py tests/test_sharded_gpu.py -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
timeout 1200 python bench.py --no-knn --no-cpu-baseline > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2o_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2o_bench.json'))
for k in ('value','ms_per_step','e2e','clocks','parity_vs_n1'): print(k, json.dumps(d[k]))
r=d['roofline']; print('roofline', r['achieved'], r['frac'], r['kernel_ms_avg'])
for k,v in d['roofline_by_data'].items(): print(k, json.dumps(v))"
timeout 300 python scripts/step_breakdown.py --images 31250 > gpurun_out/r2o_breakdown.log 2>&1; cat gpurun_out/r2o_breakdown.log
