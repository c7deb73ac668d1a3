mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu_r02c.log 2>&1; echo "pytest exit=$?"; tail -6 gpurun_out/pytest_gpu_r02c.log
timeout 400 python bench.py > gpurun_out/r02_bench_cfg1_tf32.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 exit=$?"; tail -c 600 gpurun_out/bench_cfg1.err
timeout 300 python bench.py --config 0 > gpurun_out/r02_bench_cfg0.json 2> gpurun_out/bench_cfg0.err; echo "cfg0 exit=$?"; tail -c 400 gpurun_out/bench_cfg0.err
python - <<'PY'
import json,glob
for f in ['gpurun_out/r02_bench_cfg1_tf32.json','gpurun_out/r02_bench_cfg0.json']:
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1])
    print(f, 'value %.1f ms %.2f e2e %s api %s roof %.0f/%.0f=%.3f parity %s cpu %s' % (j['value'], j['ms_per_step'], j.get('e2e',{}).get('value'), j.get('api_path'), j['roofline']['achieved'], j['roofline']['peak'], j['roofline']['frac'], j.get('parity'), j.get('cpu_baseline')))
    for k,v in j['kernels'].items():
        print('  %-28s n=%3d %8.3f ms %5.1f%% %s %s' % (k, v['launches_per_step'], v['ms_per_step'], 100*v['share'], ('%.0f TF/s' % v['tflops']) if 'tflops' in v else '', ('%.0f GB/s frac %.2f' % (v['gbs'], v['frac'])) if 'gbs' in v else ''))
PY
