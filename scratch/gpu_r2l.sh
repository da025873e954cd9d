#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "c_driver or canonicalize or fall_back or kernel_selection or too_wide" 2>&1 | grep -v "^$" | head -150 > gpurun_out/r2l_pytest.txt; tail -5 gpurun_out/r2l_pytest.txt
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2l_bench_stencil27-128.json 2> gpurun_out/r2l_bench.err; tail -c 600 gpurun_out/r2l_bench_stencil27-128.json
