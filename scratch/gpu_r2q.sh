#!/bin/bash
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2q_pytest.txt; tail -6 gpurun_out/r2q_pytest.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2q_bench_stencil27-128.json 2> gpurun_out/r2q_bench.err; tail -c 600 gpurun_out/r2q_bench_stencil27-128.json
for ch in 1 8 27 32 48; do
TSG_PLANS_CHAIN=$ch timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2q_bench_chain$ch.json 2>> gpurun_out/r2q_bench.err
done
TSG_PLANS_CTA=128 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2q_bench_cta128.json 2>> gpurun_out/r2q_bench.err
TSG_PLANS_CTA=128 TSG_PLANS_CHAIN=32 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2q_bench_cta128_chain32.json 2>> gpurun_out/r2q_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2q_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2q_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_numeric_from_plans_rows -c 1 -o gpurun_out/r2q_plans_rows python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2q_ncu2.log 2>&1
tail -5 gpurun_out/r2q_bench.err
