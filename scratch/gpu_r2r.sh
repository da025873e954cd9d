#!/bin/bash
set -x
mkdir -p gpurun_out
make -s -C driver
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2r_pytest.txt; tail -6 gpurun_out/r2r_pytest.txt
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r2r_smoke.txt 2>&1; tail -3 gpurun_out/r2r_smoke.txt
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2r_bench_stencil27-128.json 2> gpurun_out/r2r_bench.err; tail -c 600 gpurun_out/r2r_bench_stencil27-128.json
TSG_S1_KEEP_BITMAPS=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_bench_nobm.json 2>> gpurun_out/r2r_bench.err
TSG_PLANS=0 timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2r_bench_generic.json 2>> gpurun_out/r2r_bench.err
for wl in blockfem-2M lap2d-256 mixed-fem-stencil; do
timeout 400 python bench.py --workload $wl --steps 10 --warmup 3 --e2e-steps 1 > gpurun_out/r2r_bench_$wl.json 2>> gpurun_out/r2r_bench.err
done
TSG_PLANS=0 TSG_STEP3=dense timeout 400 python bench.py --workload mixed-fem-stencil --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_bench_mixed_dense.json 2>> gpurun_out/r2r_bench.err
TSG_PLANS=0 TSG_STEP3=gather timeout 400 python bench.py --workload mixed-fem-stencil --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_bench_mixed_gather.json 2>> gpurun_out/r2r_bench.err
TSG_PLANS=0 timeout 400 python bench.py --workload mixed-fem-stencil --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_bench_mixed_auto.json 2>> gpurun_out/r2r_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/r2r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_s1_fill -c 1 -o gpurun_out/r2r_s1_fill python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r2r_ncu3.log 2>&1
tail -5 gpurun_out/r2r_bench.err
