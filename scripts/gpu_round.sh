#!/bin/bash
# Full single-GPU round: parity tests, smoke, bench (ours + reference arm), ncu launch list of the bench
# command, one ncu --set full capture of K2 and of K3.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 4000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
cut -c1-300 gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 3 -c 1 -o gpurun_out/k2_full -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 > gpurun_out/ncu_k2.log 2>&1; echo "ncu k2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn3_kernel -s 1 -c 1 -o gpurun_out/k3_full -f python scripts/quick_knn.py --n 1000000 --rows 18944 --iters 1 > gpurun_out/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan1_kernel -s 3 -c 1 -o gpurun_out/k1_full -f python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --iters 3 > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 rc=$?"
ls -la gpurun_out/*.ncu-rep
