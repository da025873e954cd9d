#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python scratch/sanitize_r4.py > gpurun_out/r4i_memcheck.txt 2>&1
echo "memcheck exit $?" >> gpurun_out/r4i_memcheck.txt
grep -c "^ok" gpurun_out/r4i_memcheck.txt; tail -6 gpurun_out/r4i_memcheck.txt
