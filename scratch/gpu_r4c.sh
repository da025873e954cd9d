#!/bin/bash
# ncu of the two numeric kernels of the general-tile path on block-FEM 2M at 32x32
mkdir -p gpurun_out
TSG_GT_NUMERIC=dense timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_g_numeric_dense32 -s 3 -c 1 -o gpurun_out/r4c_dense32 python bench.py --tile 32 32 --workload blockfem-2M --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4c_ncu1.log 2>&1
TSG_GT_NUMERIC=gather timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_g_numeric -s 3 -c 1 -o gpurun_out/r4c_gather python bench.py --tile 32 32 --workload blockfem-2M --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4c_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/r4c_ncu1.log gpurun_out/r4c_ncu2.log
