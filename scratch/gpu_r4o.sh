#!/bin/bash
# one GPU, shipped build (row plans on): launch list of the default bench command and one --set full capture of the numeric kernel
mkdir -p gpurun_out
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4o_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4o_ncu1.log 2>&1; echo "ncu1 exit $?"
timeout 70 ncu --set full --clock-control none --import-source on -k regex:k_numeric_from_rowplans -s 3 -c 1 -o gpurun_out/r4o_rowplans python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4o_ncu2.log 2>&1; echo "ncu2 exit $?"
ls -la gpurun_out/ | tail -5
