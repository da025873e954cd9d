#!/bin/bash
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2y_pytest.txt; tail -6 gpurun_out/r2y_pytest.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2y_smoke.txt 2>&1; tail -3 gpurun_out/r2y_smoke.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2y_bench_stencil27-128.json 2> gpurun_out/r2y_bench.err; tail -c 600 gpurun_out/r2y_bench_stencil27-128.json
for wl in lap2d-256 stencil27-64 blockfem-2M rmat-s16-aat; do
timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2y_bench_$wl.json 2>> gpurun_out/r2y_bench.err
done
TSG_ROWPLANS_MIN_ROWS=64 timeout 400 python bench.py --workload lap2d-256 --steps 5 --warmup 3 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2y_bench_lap2d-256_templates.json 2>> gpurun_out/r2y_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2y_bench_reference.json 2>> gpurun_out/r2y_bench.err
tail -5 gpurun_out/r2y_bench.err
