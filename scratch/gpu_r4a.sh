#!/bin/bash
# round 2, session 2, call 1: general-tile path on the GPU (tests first, then two bench lines)
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gentile_gpu.py -x -q > gpurun_out/r4a_pytest_gentile.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4a_pytest_gentile.txt
tail -15 gpurun_out/r4a_pytest_gentile.txt
timeout 200 python bench.py --tile 32 32 --workload blockfem-2M --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4a_bench_blockfem-2M_32x32.json 2> gpurun_out/r4a_bench_blockfem.err
tail -c 1500 gpurun_out/r4a_bench_blockfem-2M_32x32.json; tail -3 gpurun_out/r4a_bench_blockfem.err
timeout 200 python bench.py --tile 32 32 --workload stencil27-64 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4a_bench_stencil27-64_32x32.json 2> gpurun_out/r4a_bench_stencil.err
tail -c 1200 gpurun_out/r4a_bench_stencil27-64_32x32.json; tail -3 gpurun_out/r4a_bench_stencil.err
