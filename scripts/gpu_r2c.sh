#!/bin/bash
# round 2, call C: scan GPU tests + step breakdown (PDL on) + ncu full of the merge kernel on the small shard
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_scan_gpu.py tests/test_exact_gpu.py tests/test_fullsize_gpu.py tests/test_indices_gpu.py tests/test_sharded_gpu.py -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -8 gpurun_out/r2c_pytest.log
timeout 300 python scripts/step_breakdown.py > gpurun_out/r2c_breakdown.log 2>&1; echo "breakdown rc=$?"
cat gpurun_out/r2c_breakdown.log
timeout 300 python scripts/step_breakdown.py --dim 768 --images 312500 --iters 100 > gpurun_out/r2c_breakdown768.log 2>&1
cat gpurun_out/r2c_breakdown768.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2c_launches_small.csv python scripts/step_breakdown.py --images 31250 --iters 3 > gpurun_out/r2c_ncu.log 2>&1; echo "ncu rc=$?"
grep -v "^==" gpurun_out/r2c_launches_small.csv | awk -F'","' '{print $5, $NF}' | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:merge_topk_kernel -s 12 -c 1 -o gpurun_out/r2c_merge_full -f python scripts/step_breakdown.py --images 31250 --iters 3 > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu merge rc=$?"
