#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r3d_pytest.txt; tail -3 gpurun_out/r3d_pytest.txt
timeout 300 python scratch/time_tile2csr.py stencil27-128 > gpurun_out/r3d_tile2csr_stencil27-128.json 2> gpurun_out/r3d.err; cat gpurun_out/r3d_tile2csr_stencil27-128.json
timeout 300 python scratch/time_tile2csr.py blockfem-2M > gpurun_out/r3d_tile2csr_blockfem-2M.json 2>> gpurun_out/r3d.err; cat gpurun_out/r3d_tile2csr_blockfem-2M.json
timeout 600 python bench.py --workload rmat-s18-aat --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r3d_bench_rmat-s18-aat.json 2>> gpurun_out/r3d.err; tail -c 400 gpurun_out/r3d_bench_rmat-s18-aat.json
tail -3 gpurun_out/r3d.err
