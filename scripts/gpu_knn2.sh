#!/bin/bash
echo "1cta N256 no-epilogue"; SSW_KNN_1CTA=1 SSW_KNN_N256=1 SSW_KNN_DEBUG=1 timeout 120 python scripts/quick_knn.py --n 1000000 --rows 37888 --iters 2 2>&1 | tail -2 | head -1
echo "1cta N256 with epilogue (serialized)"; SSW_KNN_1CTA=1 SSW_KNN_N256=1 timeout 120 python scripts/quick_knn.py --n 1000000 --rows 37888 --iters 2 2>&1 | tail -2
echo "1cta N128 no-epilogue"; SSW_KNN_1CTA=1 SSW_KNN_DEBUG=1 timeout 120 python scripts/quick_knn.py --n 1000000 --rows 37888 --iters 2 2>&1 | tail -2 | head -1
