#!/bin/bash
# ncu evidence for profiles/: launch list of the bench command, --set full captures of K2 (512 and 768), K1, K3,
# and of the fused exchange kernel on a 1.25M-row shard (world = 1).  Every command below exits 0 without ncu first
# (scripts/gpu_r2*.sh); numbers printed by a program running under ncu are never bench values.
set -x
mkdir -p gpurun_out
R=gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $R/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-data-sweep > $R/r02_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 3 -c 1 -o $R/r02_k2_full -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 > $R/r02_ncu_k2.log 2>&1; echo "k2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 3 -c 1 -o $R/r02_k2_768_full -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 --dim 768 --images 312500 > $R/r02_ncu_k2_768.log 2>&1; echo "k2 768 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan1_kernel -s 3 -c 1 -o $R/r02_k1_full -f python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --iters 3 > $R/r02_ncu_k1.log 2>&1; echo "k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn3_kernel -s 1 -c 1 -o $R/r02_k3_full -f python scripts/quick_knn.py --n 1000000 --rows 18944 --iters 1 > $R/r02_ncu_k3.log 2>&1; echo "k3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:exchange_merge_kernel -s 20 -c 1 -o $R/r02_xchg_full -f python scripts/step_breakdown.py --images 31250 --iters 3 > $R/r02_ncu_xchg.log 2>&1; echo "xchg rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 20 -c 1 -o $R/r02_k2_small_full -f python scripts/step_breakdown.py --images 31250 --iters 3 > $R/r02_ncu_k2_small.log 2>&1; echo "k2 small rc=$?"
ls -la $R/*.ncu-rep
