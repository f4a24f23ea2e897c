#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/l_tiny.csv python scripts/quick_scan.py --nq 64 --k 50 --mode 2 --excl 50 --images 500 --iters 8 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/l_tiny1.csv python scripts/quick_scan.py --nq 1 --k 50 --mode 1 --excl 50 --images 500 --iters 8 > /dev/null 2>&1
