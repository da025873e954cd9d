#!/bin/bash
# round 2: full GPU suite on 2 GPUs (incl. the multi-rank test), default bench line, 2-GPU bench lines (both broadcast modes)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2g_pytest.txt; tail -5 gpurun_out/r2g_pytest.txt
timeout 400 python bench.py > gpurun_out/r2g_bench_stencil27-128.json 2> gpurun_out/r2g_bench.err; tail -c 2500 gpurun_out/r2g_bench_stencil27-128.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2g_bench_stencil27-128_2gpu.json 2>> gpurun_out/r2g_bench.err; tail -c 1500 gpurun_out/r2g_bench_stencil27-128_2gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --bcast tiled > gpurun_out/r2g_bench_stencil27-128_2gpu_tiled.json 2>> gpurun_out/r2g_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2g_bench_ref_2gpu.json 2>> gpurun_out/r2g_bench.err
tail -5 gpurun_out/r2g_bench.err
ls gpurun_out | grep nccl | head
