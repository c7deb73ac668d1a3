#!/bin/bash
# timing experiments: per-kernel ms of the conv family under SZN_DBG switches (SZN_DBG_MODE selects fwd/dgrad 0 or wgrad 2)
for m in 0 2; do for d in 0 1 2 3; do
  SZN_DBG_MODE=$m SZN_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/dbg$m$d.log 2>&1
  python - <<PY
import json
l=[x for x in open('gpurun_out/dbg$m$d.log') if x.startswith('{')]
j=json.loads(l[-1])
print('mode=$m dbg=$d', {k:round(v['ms_per_step'],2) for k,v in j['kernels'].items() if k in ('szn_conv_fwd','szn_conv_dgrad','szn_conv_wgrad')}, 'step', round(j['ms_per_step'],1))
PY
done; done
