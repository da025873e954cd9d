#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r3g_$name.json 2>> gpurun_out/r3g.err
  python - "$name" <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r3g_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('sweep', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), round(d['config']['steps_ms']['step3'],3), d['parity']['rowsums_equal'])
PY
}
run base A=1
run round4 TSG_PLANS_ROUND4=1
run round4_chain16 TSG_PLANS_ROUND4=1 TSG_PLANS_CHAIN=16
run round4_chain8 TSG_PLANS_ROUND4=1 TSG_PLANS_CHAIN=8
tail -2 gpurun_out/r3g.err
