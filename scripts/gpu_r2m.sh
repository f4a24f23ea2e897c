#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_service_gpu.py -m gpu -q -x -s 2>&1 | tail -3
timeout 900 python scripts/bench_service.py > gpurun_out/r2m_service.json 2> gpurun_out/r2m_service.err; echo "service rc=$?"; tail -6 gpurun_out/r2m_service.err
