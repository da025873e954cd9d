#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gentile_gpu.py -x -q -k "numeric_kernels or dense32 or larger" > gpurun_out/r4e_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4e_pytest.txt
tail -5 gpurun_out/r4e_pytest.txt
TSG_GT_NUMERIC=dense timeout 200 python bench.py --tile 32 32 --workload blockfem-2M --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4e_blockfem.json 2>> gpurun_out/r4e.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r4e_blockfem.json').read().strip().splitlines()[-1])
print('r4e', round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['parity']['rowsums_equal'], d['parity']['rowcounts_equal'])
PY
TSG_GT_NUMERIC=dense timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_g_numeric_dense32 -s 3 -c 1 -o gpurun_out/r4e_dense32 python bench.py --tile 32 32 --workload blockfem-2M --steps 1 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4e_ncu.log 2>&1
tail -n 3 gpurun_out/r4e.err
