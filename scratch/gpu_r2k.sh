#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2k_pytest.txt; tail -5 gpurun_out/r2k_pytest.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2k_bench_stencil27-128.json 2> gpurun_out/r2k_bench.err; tail -c 1200 gpurun_out/r2k_bench_stencil27-128.json
timeout 300 python bench.py --workload blockfem-2M --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r2k_bench_blockfem.json 2>> gpurun_out/r2k_bench.err
timeout 300 python bench.py --workload lap2d-256 --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r2k_bench_lap2d.json 2>> gpurun_out/r2k_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2k_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_numeric_from_plans -c 1 -o gpurun_out/r2k_plans python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2k_ncu2.log 2>&1
tail -5 gpurun_out/r2k_bench.err
