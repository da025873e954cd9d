#!/bin/bash
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r2t_pytest.txt; tail -6 gpurun_out/r2t_pytest.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2t_bench_stencil27-128.json 2> gpurun_out/r2t_bench.err; tail -c 600 gpurun_out/r2t_bench_stencil27-128.json
TSG_PLANS=2 timeout 400 python bench.py --workload blockfem-2M --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2t_bench_blockfem-2M_plans.json 2>> gpurun_out/r2t_bench.err
TSG_PLANS=2 timeout 400 python bench.py --workload mixed-fem-stencil --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2t_bench_mixed_plans.json 2>> gpurun_out/r2t_bench.err
timeout 400 python bench.py --workload stencil27-64 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2t_bench_stencil27-64.json 2>> gpurun_out/r2t_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2t_ncu1.log 2>&1
tail -5 gpurun_out/r2t_bench.err
