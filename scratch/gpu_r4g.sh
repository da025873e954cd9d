#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/r4g_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/r4g_smoke.txt; tail -12 gpurun_out/r4g_smoke.txt
timeout 300 python bench.py --workload stencil27-64 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4g_bench_stencil27-64.json 2>> gpurun_out/r4g.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r4g_launches_stencil27-64.csv python bench.py --workload stencil27-64 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4g_ncu.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4g_bench_stencil27-64.json').read().strip().splitlines()[-1])
print('r4g', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['gpu_launches'], d['parity']['rowsums_equal'])
PY
timeout 300 python bench.py --tile 32 32 --workload stencil27-128 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4g_bench_stencil27-128_32x32.json 2>> gpurun_out/r4g.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4g_bench_stencil27-128_32x32.json').read().strip().splitlines()[-1])
print('r4g 128^3 32x32', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['e2e']['ms_per_step'], d['parity'])
PY
tail -n 3 gpurun_out/r4g.err
