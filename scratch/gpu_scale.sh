#!/bin/bash
# scaling runs: bash scratch/gpu_scale.sh N TAG [rmat]
N=$1; TAG=$2
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
[ "$N" = "1" ] && RUN="python"
set -x
make -s -C driver
timeout 600 $RUN $( [ "$N" != "1" ] && echo --master-port 29521 ) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_stencil27-128_${N}gpu.json 2> gpurun_out/${TAG}_scale_${N}.err
tail -c 400 gpurun_out/${TAG}_bench_stencil27-128_${N}gpu.json
if [ "$3" = "rmat" ]; then
  timeout 2400 $RUN $( [ "$N" != "1" ] && echo --master-port 29522 ) bench.py --gpus $N --workload rmat-s24 --steps 1 --warmup 3 --e2e-steps 1 --parity-counts > gpurun_out/${TAG}_bench_rmat-s24_${N}gpu.json 2>> gpurun_out/${TAG}_scale_${N}.err
  tail -c 600 gpurun_out/${TAG}_bench_rmat-s24_${N}gpu.json
fi
tail -3 gpurun_out/${TAG}_scale_${N}.err
