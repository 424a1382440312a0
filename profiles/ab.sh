#!/bin/bash
# usage: profiles/ab.sh ENVVAR v1 v2 ... ; prints ms/step and per-kernel ms for each value
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > /tmp/ab.json 2>/tmp/ab.err || { echo "$var=$v FAILED"; tail -3 /tmp/ab.err; continue; }
  python - "$var=$v" <<'PY'
import json,sys
d=json.load(open('/tmp/ab.json'))
print(sys.argv[1], 'ms/step %.4f'%d['ms_per_step'], ' '.join('%s=%.4f'%(k,v['ms_avg']) for k,v in d['kernels'].items() if v['ms_avg']>0.05))
PY
done
