#!/bin/bash
# 2-GPU box: the sharded tests (NCCL + fused exchange + pipelined steps in every exchange shape)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -x -q 2>&1 | tail -5
