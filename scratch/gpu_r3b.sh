#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r3b_bench_stencil27-128.json 2> gpurun_out/r3b_bench.err; tail -c 600 gpurun_out/r3b_bench_stencil27-128.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3b_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_numeric_from_plans_rows -c 1 -o gpurun_out/r3b_plans_rows python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3b_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_rows_instantiate -c 1 -o gpurun_out/r3b_rows_instantiate python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3b_ncu3.log 2>&1
tail -3 gpurun_out/r3b_bench.err
