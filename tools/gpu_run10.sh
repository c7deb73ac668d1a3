mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_edge_cases_gpu.py -q -x --timeout 200 -k "conv1_1 or golden or ragged or fp32_grade" 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 20 --no-e2e > gpurun_out/r02_bench_cfg1_c.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 exit=$?"; tail -c 300 gpurun_out/bench_cfg1.err
python - <<'PY'
import json
for f in ['gpurun_out/r02_bench_cfg1_c.json']:
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1])
    print(f, 'value %.1f ms %.2f' % (j['value'], j['ms_per_step']))
    for k,v in list(j['kernels'].items()):
        print('  %-28s n=%3d %8.3f ms %5.1f%% %s %s' % (k, v['launches_per_step'], v['ms_per_step'], 100*v['share'], ('%.0f TF/s' % v['tflops']) if 'tflops' in v else '', ('%.0f GB/s %.2f' % (v['gbs'], v['frac'])) if 'gbs' in v else ''))
PY
