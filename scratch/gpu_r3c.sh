#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r3c_pytest.txt; tail -3 gpurun_out/r3c_pytest.txt
timeout 300 python scratch/time_tile2csr.py stencil27-128 > gpurun_out/r3c_tile2csr_stencil27-128.json 2> gpurun_out/r3c.err; cat gpurun_out/r3c_tile2csr_stencil27-128.json
timeout 300 python scratch/time_tile2csr.py blockfem-2M > gpurun_out/r3c_tile2csr_blockfem-2M.json 2>> gpurun_out/r3c.err; cat gpurun_out/r3c_tile2csr_blockfem-2M.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r3c_bench_stencil27-128.json 2>> gpurun_out/r3c.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tile2csr_rows -c 1 -o gpurun_out/r3c_tile2csr python scratch/time_tile2csr.py stencil27-128 > gpurun_out/r3c_ncu.log 2>&1
tail -3 gpurun_out/r3c.err
