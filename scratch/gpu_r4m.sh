#!/bin/bash
# final evidence of the shipped build (row plans on), one GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r4m_pytest_gpu.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r4m_pytest_gpu.txt; tail -4 gpurun_out/r4m_pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/r4m_smoke.txt 2>&1; echo "smoke exit $?" >> gpurun_out/r4m_smoke.txt; tail -2 gpurun_out/r4m_smoke.txt
timeout 600 python bench.py > gpurun_out/r4m_bench_stencil27-128.json 2> gpurun_out/r4m_bench.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4m_bench_stencil27-128.json').read().strip().splitlines()[-1])
print('r4m', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['gpu_launches'], d['parity'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['ms_per_step'], d.get('cpu_baseline',{}).get('value'))
PY
tail -n 3 gpurun_out/r4m_bench.err
