#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_exact_gpu.py tests/test_fullsize_gpu.py tests/test_indices_gpu.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -5 gpurun_out/r2h_pytest.log
timeout 300 python scripts/step_breakdown.py --images 31250 62500 250000 > gpurun_out/r2h_breakdown.log 2>&1; cat gpurun_out/r2h_breakdown.log
timeout 300 python scripts/step_breakdown.py --dim 768 --images 312500 --iters 100 2>&1 | head -4
timeout 1200 python bench.py --no-knn --no-cpu-baseline > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2h_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json'))
for k in ('value','ms_per_step','e2e','roofline','clocks','parity_vs_n1','single_query'): print(k, json.dumps(d[k]))
for k,v in d['roofline_by_data'].items(): print(k, json.dumps(v))
print(json.dumps(d['config2_multiscale_120k_images']))"
