#!/bin/bash
mkdir -p gpurun_out
timeout 120 python scripts/ab_pipeline.py > gpurun_out/r2w_ab_small.log 2>&1; cat gpurun_out/r2w_ab_small.log
timeout 120 python scripts/ab_pipeline.py --images 62500 --shapes inline 0 4 --rounds 3 > gpurun_out/r2w_ab_n4.log 2>&1; cat gpurun_out/r2w_ab_n4.log
