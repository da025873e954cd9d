#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "plans or templates or config2 or spgemm_matches or mixed" > gpurun_out/r4k_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4k_pytest.txt
tail -4 gpurun_out/r4k_pytest.txt
run() { name=$1; wl=$2; shift 2
  env "$@" timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4k_$name.json 2>> gpurun_out/r4k.err
  python - $name <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r4k_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('r4k', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['gpu_launches'], d['parity']['rowsums_equal'], d['parity'].get('rowcounts_equal'))
PY
}
run rowplan_128 stencil27-128 X=1
run norowplan_128 stencil27-128 TSG_ROWPLAN_NUMERIC=0
run rowplan_64 stencil27-64 X=1
tail -n 3 gpurun_out/r4k.err
