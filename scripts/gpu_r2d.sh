#!/bin/bash
# round 2, call D: scan tests, breakdown, full default bench + reference arm
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_exact_gpu.py tests/test_sharded_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
timeout 300 python scripts/step_breakdown.py --images 31250 250000 > gpurun_out/r2d_breakdown.log 2>&1; echo "breakdown rc=$?"
cat gpurun_out/r2d_breakdown.log
( time timeout 1200 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err ) 2>&1 | tail -4; echo "bench rc=$?"
tail -5 gpurun_out/r2d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2d_bench.json'))
print(json.dumps({k:d[k] for k in d if k not in ('config',)}, indent=1)[:9000])"
( time timeout 600 python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/r2d_bench_ref.json 2> gpurun_out/r2d_bench_ref.err ) 2>&1 | tail -4
cut -c1-1500 gpurun_out/r2d_bench_ref.json
