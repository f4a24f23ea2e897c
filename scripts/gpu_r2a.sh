#!/bin/bash
# round 2, call A: full GPU test-suite + step breakdown on small shards
set -x
mkdir -p gpurun_out
nvidia-smi -L; free -g | head -2; nproc
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
grep -E "exact mode|fp16 storage alone|float32 storage" gpurun_out/r2a_pytest.log
timeout 300 python scripts/step_breakdown.py > gpurun_out/r2a_breakdown.log 2>&1; echo "breakdown rc=$?"
cat gpurun_out/r2a_breakdown.log
