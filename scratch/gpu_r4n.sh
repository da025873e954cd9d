#!/bin/bash
# 2 GPUs, shipped build (row plans on): the config-2 line with parity, as the driver's scaling run launches it
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4n_bench_stencil27-128_2gpu.json 2> gpurun_out/r4n_bench.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r4n_bench_stencil27-128_2gpu.json
tail -n 5 gpurun_out/r4n_bench.err
