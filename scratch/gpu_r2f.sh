#!/bin/bash
# round 2, first GPU pass: parity of the new numeric path, A/B against the gather, launch list + full ncu capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2f_pytest.txt; tail -5 gpurun_out/r2f_pytest.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench_stencil27-128.json 2> gpurun_out/r2f_bench.err; tail -c 1500 gpurun_out/r2f_bench_stencil27-128.json
timeout 300 python bench.py --workload blockfem-2M --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_bench_blockfem.json 2>> gpurun_out/r2f_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step3_rows -c 1 -o gpurun_out/r2f_rows python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2f_ncu2.log 2>&1
ls -la gpurun_out | tail -20
