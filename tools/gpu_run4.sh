mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_edge_cases_gpu.py tests/test_trainer_gpu.py tests/test_head_gpu.py tests/test_train_step_gpu.py -q -x --timeout 200 2>&1 | tail -8
timeout 400 python bench.py > gpurun_out/r02_bench_cfg1_tf32.json 2> gpurun_out/bench_cfg1.err; echo "cfg1 exit=$?"; tail -c 600 gpurun_out/bench_cfg1.err
timeout 300 python bench.py --precision fp32 --no-cpu-baseline > gpurun_out/r02_bench_cfg1_fp32.json 2> gpurun_out/bench_cfg1_fp32.err; echo "cfg1 fp32 exit=$?"; tail -c 400 gpurun_out/bench_cfg1_fp32.err
timeout 300 python bench.py --config 0 > gpurun_out/r02_bench_cfg0.json 2> gpurun_out/bench_cfg0.err; echo "cfg0 exit=$?"; tail -c 400 gpurun_out/bench_cfg0.err
timeout 300 python bench.py --config 3 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 exit=$?"; tail -c 400 gpurun_out/bench_cfg3.err
timeout 300 python bench.py --config 2 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit=$?"; tail -c 400 gpurun_out/bench_cfg2.err
timeout 300 python bench.py --config 4 --steps 20 --no-cpu-baseline > gpurun_out/r02_bench_cfg4.json 2> gpurun_out/bench_cfg4.err; echo "cfg4 exit=$?"; tail -c 400 gpurun_out/bench_cfg4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_bench_cfg*.json')):
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1])
    print(f, 'value %.1f ms %.2f e2e %s roof %.0f/%.0f=%.3f parity %s cpu %s' % (j['value'], j['ms_per_step'], j.get('e2e',{}).get('value'), j['roofline']['achieved'], j['roofline']['peak'], j['roofline']['frac'], j.get('parity'), j.get('cpu_baseline',{}).get('value')))
    if 'cfg1_tf32' in f:
        for k,v in j['kernels'].items():
            print('  %-28s n=%3d %8.3f ms %5.1f%% %s %s' % (k, v['launches_per_step'], v['ms_per_step'], 100*v['share'], ('%.0f TF/s' % v['tflops']) if 'tflops' in v else '', ('%.0f GB/s frac %.2f' % (v['gbs'], v['frac'])) if 'gbs' in v else ''))
PY
