#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "canonicalize or out_of_range or row_slice or c_driver" 2>&1 | tail -15 > gpurun_out/r2i_pytest.txt; tail -3 gpurun_out/r2i_pytest.txt
timeout 900 python bench.py --workload rmat-s22 --steps 1 --warmup 3 --e2e-steps 1 --parity-counts > gpurun_out/r2i_bench_rmat-s22.json 2> gpurun_out/r2i_bench.err; tail -c 1800 gpurun_out/r2i_bench_rmat-s22.json
timeout 1500 python bench.py --workload rmat-s20-aat --steps 1 --warmup 3 --e2e-steps 1 --parity-counts > gpurun_out/r2i_bench_rmat-s20-aat.json 2>> gpurun_out/r2i_bench.err; tail -c 1800 gpurun_out/r2i_bench_rmat-s20-aat.json
tail -5 gpurun_out/r2i_bench.err
