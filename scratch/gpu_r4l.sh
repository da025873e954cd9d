#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r4l_$name.json 2>> gpurun_out/r4l.err
  python - $name <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r4l_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('r4l', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['config']['steps_ms']['step3'],3), round(d['config']['steps_ms']['step2'],3))
PY
}
run carve22 TSG_PLANS_CARVEOUT=22
run carve32 TSG_PLANS_CARVEOUT=32
run carve40 TSG_PLANS_CARVEOUT=40
run chain10 TSG_PLANS_CHAIN=10
run chain16 TSG_PLANS_CHAIN=16
run chain24 TSG_PLANS_CHAIN=24
tail -n 2 gpurun_out/r4l.err
