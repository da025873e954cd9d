#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --e2e-steps 1 > gpurun_out/r3e_$name.json 2>> gpurun_out/r3e.err
  python - "$name" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3e_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('sweep', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['config']['steps_ms']['step3'],3))
PY
}
run base A=1
run chain12 TSG_PLANS_CHAIN=12
run chain20 TSG_PLANS_CHAIN=20
run chain24 TSG_PLANS_CHAIN=24
run carve26 TSG_PLANS_CARVEOUT=26
run carve38 TSG_PLANS_CARVEOUT=38
run carve44 TSG_PLANS_CARVEOUT=44
tail -2 gpurun_out/r3e.err
