#!/bin/bash
set -x
mkdir -p gpurun_out
for cv in -1 default 20 50 100; do
  if [ "$cv" = "default" ]; then unset TSG_PLANS_CARVEOUT; else export TSG_PLANS_CARVEOUT=$cv; fi
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3a_bench_carveout_$cv.json 2>> gpurun_out/r3a_bench.err
  python - "$cv" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3a_bench_carveout_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('carveout', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms']['step3'])
PY
done
tail -3 gpurun_out/r3a_bench.err
