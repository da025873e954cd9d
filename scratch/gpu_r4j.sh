#!/bin/bash
# final evidence of the shipped build, one GPU: all GPU tests, smoke, the default bench line (cpu baseline, parity), the reference arm, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4j_pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r4j_pytest_gpu.txt; tail -4 gpurun_out/r4j_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r4j_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/r4j_smoke.txt; tail -2 gpurun_out/r4j_smoke.txt
timeout 600 python bench.py > gpurun_out/r4j_bench_stencil27-128.json 2> gpurun_out/r4j_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r4j_bench_reference.json 2>> gpurun_out/r4j_bench.err; echo "ref exit $?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4j_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4j_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4j_bench_stencil27-128.json').read().strip().splitlines()[-1])
print('r4j', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['gpu_launches'], d['parity'], d['roofline']['frac'], d['e2e']['ms_per_step'], d.get('cpu_baseline',{}).get('value'))
r=json.loads(open('gpurun_out/r4j_bench_reference.json').read().strip().splitlines()[-1])
print('ref', r['value'], r['cpu_baseline']['cores'])
PY
tail -n 3 gpurun_out/r4j_bench.err
