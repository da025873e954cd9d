#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r3f_pytest.txt; tail -4 gpurun_out/r3f_pytest.txt
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3f_$name.json 2>> gpurun_out/r3f.err
  python - "$name" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3f_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('sweep', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['config']['steps_ms']['step3'],3))
PY
}
run chain12_default A=1
run chain10 TSG_PLANS_CHAIN=10
run chain14 TSG_PLANS_CHAIN=14
run chain8 TSG_PLANS_CHAIN=8
timeout 300 python scratch/time_tile2csr.py blockfem-2M > gpurun_out/r3f_tile2csr_blockfem-2M.json 2>> gpurun_out/r3f.err; cat gpurun_out/r3f_tile2csr_blockfem-2M.json
tail -2 gpurun_out/r3f.err
