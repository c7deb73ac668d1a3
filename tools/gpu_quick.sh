#!/bin/bash
# quick GPU check: kernel parity + bench kernel table
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_kernels_gpu.py tests/test_head_gpu.py -q -x --timeout 60 2>&1 | tail -3
timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_quick.log') if x.startswith('{')]
if not l:
    print(open('gpurun_out/bench_quick.log').read()[-2000:])
else:
    j=json.loads(l[-1])
    print('value', round(j['value'],2), 'ms/step', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), 'roofline', round(j['roofline']['achieved'],1), round(j['roofline']['frac'],3))
    for k,v in j['kernels'].items():
        print('  %-28s n=%3d %8.3f ms %5.1f%% %s' % (k, v['launches_per_step'], v['ms_per_step'], 100*v['share'], ('%.0f TF/s' % v['tflops']) if 'tflops' in v else ''))
PY
