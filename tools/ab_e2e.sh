#!/bin/bash
# A/B of the multi-threaded e2e leg: tools/ab_e2e.sh "<env> -- <bench args>" ...
for v in "$@"; do
  envs="${v%% -- *}"; args="${v##* -- }"
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-steady-state $args > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json,sys
try:
    d=json.load(open('gpurun_out/ab.json')); e=d['e2e']; print(sys.argv[1], '| value', round(d['value']), '| e2e', round(e['value']), e.get('host_threads'), round(e['ms_per_step'],2), {k:round(v['value']) for k,v in e.items() if isinstance(v,dict)})
except Exception as ex: print(sys.argv[1], 'failed', ex, open('gpurun_out/ab.err').read()[-300:])
" "$v"
done
