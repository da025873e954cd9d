#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "plans or templates or golden or config2 or config1 or spgemm_matches" > gpurun_out/r4h_pytest.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r4h_pytest.txt
tail -4 gpurun_out/r4h_pytest.txt
for wl in stencil27-64 stencil27-128; do
timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/r4h_bench_$wl.json 2>> gpurun_out/r4h.err
python - $wl <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r4h_bench_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print('r4h', sys.argv[1], round(d['value'],1), round(d['ms_per_step'],3), d['config']['steps_ms'], d['gpu_launches'], d['parity']['rowsums_equal'], d['parity'].get('rowcounts_equal'))
PY
done
tail -n 3 gpurun_out/r4h.err
