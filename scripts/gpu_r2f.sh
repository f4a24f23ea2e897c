#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/trace_k2.py 31250 2>&1 | tail -16
timeout 900 python -m pytest tests/test_scan_gpu.py tests/test_exact_gpu.py tests/test_fullsize_gpu.py tests/test_knn_gpu.py -m gpu -q -x > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
timeout 300 python scripts/step_breakdown.py --images 31250 250000 > gpurun_out/r2f_breakdown.log 2>&1; cat gpurun_out/r2f_breakdown.log
timeout 300 python scripts/step_breakdown.py --dim 768 --images 312500 --iters 100 2>&1 | head -3
