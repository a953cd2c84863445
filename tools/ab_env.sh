#!/bin/bash
# A/B of environment switches: tools/ab_env.sh "ZB_X=1" "ZB_X=2 ZB_Y=3" ...  (one short bench.py run per setting)
for v in "$@"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-steady-state --e2e-threads 1 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/ab.json')); e=d['e2e']; print(sys.argv[1], round(d['value']), round(d['ms_per_step'],3), 'e2e', round(e['value']), {k:v['ms'] for k,v in d['kernels'].items()})
except Exception as ex: print(sys.argv[1], 'failed', ex, open('gpurun_out/ab.err').read()[-300:])
" "$v"
done
