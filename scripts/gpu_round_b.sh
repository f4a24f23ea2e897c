#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 0 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 64 --k 1 --mode 2 --excl 0 2>&1 | tail -2
timeout 200 python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 3 -c 1 -o gpurun_out/k2_r1e -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 > gpurun_out/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_kernel -s 1 -c 1 -o gpurun_out/k3_r1a -f python scripts/quick_knn.py --n 65536 --iters 1 > gpurun_out/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
