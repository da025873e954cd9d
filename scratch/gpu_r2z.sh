#!/bin/bash
# 2-GPU box: the multi-GPU parity tests and the config-2 scaling line
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q 2>&1 | tail -30 > gpurun_out/r2z_pytest_2gpu.txt; tail -5 gpurun_out/r2z_pytest_2gpu.txt
bash scratch/gpu_scale.sh 2 r2z
