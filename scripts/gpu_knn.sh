#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_knn_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 120 python scripts/quick_knn.py --n 65536 --iters 3 2>&1 | tail -2
timeout 120 python scripts/quick_knn.py --n 1000000 --rows 37888 --iters 2 2>&1 | tail -2
SSW_KNN_DEBUG=1 timeout 120 python scripts/quick_knn.py --n 1000000 --rows 37888 --iters 2 2>&1 | tail -2 | head -1
timeout 120 python scripts/quick_knn.py --n 1000000 --rows 151552 --iters 3 2>&1 | tail -2
SSW_KNN_TS=1 timeout 120 python scripts/quick_knn.py --n 1000000 --rows 151552 --iters 3 2>&1 | tail -2
