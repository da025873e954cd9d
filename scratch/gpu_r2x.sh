#!/bin/bash
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "tile_row_templates or recipe_plans" 2>&1 | tail -30 > gpurun_out/r2x_pytest_rowplans.txt; tail -8 gpurun_out/r2x_pytest_rowplans.txt
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2x_pytest.txt; tail -6 gpurun_out/r2x_pytest.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2x_bench_stencil27-128.json 2> gpurun_out/r2x_bench.err; tail -c 600 gpurun_out/r2x_bench_stencil27-128.json
TSG_ROWPLANS=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2x_bench_norowplans.json 2>> gpurun_out/r2x_bench.err
timeout 400 python bench.py --workload mixed-fem-stencil --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2x_bench_mixed.json 2>> gpurun_out/r2x_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2x_ncu1.log 2>&1
tail -5 gpurun_out/r2x_bench.err
