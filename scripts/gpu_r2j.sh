#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -s > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -12 gpurun_out/r2j_pytest.log; grep -E "float32 kNN|plain fp16|stage-1 scans of|exact mode|fp16 storage alone" gpurun_out/r2j_pytest.log
timeout 900 python scripts/bench_service.py > gpurun_out/r2j_service.json 2> gpurun_out/r2j_service.err; echo "service rc=$?"; tail -6 gpurun_out/r2j_service.err
