mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -6
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_cfg1_d.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 exit=$?"; tail -c 300 gpurun_out/bench_cfg1.err
timeout 300 python bench.py --no-cpu-baseline --config 2 --steps 20 > gpurun_out/r02_bench_cfg2_d.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit=$?"
python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_cfg1_d.json','gpurun_out/r02_bench_cfg2_d.json']:
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1])
    print(f, 'value %.1f ms %.2f e2e %s api %s roof %.0f/%.0f=%.3f' % (j['value'], j['ms_per_step'], j.get('e2e',{}).get('value'), j.get('api_path',{}).get('ms_per_step'), j['roofline']['achieved'], j['roofline']['peak'], j['roofline']['frac']))
    for k,v in list(j['kernels'].items())[:4]:
        print('  %-28s n=%3d %8.3f ms %5.1f%% %s' % (k, v['launches_per_step'], v['ms_per_step'], 100*v['share'], ('%.0f TF/s' % v['tflops']) if 'tflops' in v else ''))
PY
