#!/bin/bash
# dense32 numeric kernel of the general-tile path: tests, then block-FEM / stencil at 32x32 with both numeric kernels
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gentile_gpu.py -x -q > gpurun_out/r4d_pytest_gentile.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4d_pytest_gentile.txt
tail -12 gpurun_out/r4d_pytest_gentile.txt
run() { name=$1; wl=$2; mode=$3
  TSG_GT_NUMERIC=$mode timeout 200 python bench.py --tile 32 32 --workload $wl --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4d_$name.json 2>> gpurun_out/r4d.err
  python - "$name" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r4d_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('r4d', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['parity']['rowsums_equal'], d['parity']['rowcounts_equal'])
PY
}
run blockfem-2M_32x32_dense blockfem-2M dense
run stencil27-64_32x32_dense stencil27-64 dense

tail -3 gpurun_out/r4d.err
