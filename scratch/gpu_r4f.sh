#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r4f_pytest_parity.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4f_pytest_parity.txt
tail -5 gpurun_out/r4f_pytest_parity.txt
timeout 300 python bench.py --workload blockfem-2M --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4f_bench_blockfem-2M.json 2>> gpurun_out/r4f.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4f_bench_blockfem-2M.json').read().strip().splitlines()[-1])
print('r4f', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['parity'], d['roofline']['kernel'])
PY
tail -n 3 gpurun_out/r4f.err
