#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_service_gpu.py tests/test_knn_gpu.py tests/test_lp_gpu.py tests/test_scan_gpu.py -m gpu -q -x -s > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -6 gpurun_out/r2l_pytest.log
timeout 900 python scripts/bench_service.py > gpurun_out/r2l_service.json 2> gpurun_out/r2l_service.err; echo "service rc=$?"; tail -6 gpurun_out/r2l_service.err
