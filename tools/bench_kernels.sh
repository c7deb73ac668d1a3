#!/bin/bash
# timing experiments: per-kernel ms of the conv family under SZN_DBG switches
for d in 0 1 2 3; do
  SZN_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/dbg$d.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/dbg$d.log') if x.startswith('{')]
j=json.loads(l[-1])
print('dbg=$d', {k:round(v['ms_per_step'],2) for k,v in j['kernels'].items() if 'conv_' in k}, 'step', round(j['ms_per_step'],1))
PY
done
