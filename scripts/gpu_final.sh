#!/bin/bash
set -x
bash scripts/gpu_round.sh
R=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $R/r02_launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-data-sweep > $R/r02_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 3 -c 1 -o $R/r02_k2_full -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 > $R/r02_ncu_k2.log 2>&1; echo "k2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 3 -c 1 -o $R/r02_k2_768_full -f python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --iters 3 --dim 768 --images 312500 > $R/r02_ncu_k2_768.log 2>&1; echo "k2 768 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc8_kernel -s 20 -c 1 -o $R/r02_k2_small_full -f python scripts/step_breakdown.py --images 31250 --iters 3 > $R/r02_ncu_k2_small.log 2>&1; echo "k2 small rc=$?"
