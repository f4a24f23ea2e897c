#!/bin/bash
# round 2, call B: GPU test-suite + step breakdown (512 and 768) + ncu launch list of the small-shard step
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -25 gpurun_out/r2b_pytest.log
timeout 300 python scripts/step_breakdown.py > gpurun_out/r2b_breakdown.log 2>&1; echo "breakdown rc=$?"
cat gpurun_out/r2b_breakdown.log
timeout 300 python scripts/step_breakdown.py --dim 768 --images 39062 312500 --iters 100 > gpurun_out/r2b_breakdown768.log 2>&1; echo "breakdown768 rc=$?"
cat gpurun_out/r2b_breakdown768.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2b_launches_small.csv python scripts/step_breakdown.py --images 31250 --iters 3 > gpurun_out/r2b_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2b_launches_small.csv | awk -F'","' '{print $5, $NF}' | tail -40
