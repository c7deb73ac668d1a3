mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
timeout 300 python -m pytest tests/test_ddp_gpu.py -q --timeout 200 -s 2>&1 | tail -4
timeout 300 python bench.py --no-cpu-baseline --steps 30 --no-e2e > gpurun_out/sc_n1.json 2> gpurun_out/sc.err
timeout 300 $R 29511 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/sc_n2_bucket.json 2>> gpurun_out/sc.err; echo "exit=$?"; tail -c 600 gpurun_out/sc.err
python - <<'PY'
import json,glob
for f in ['sc_n1','sc_n2_bucket']:
    l=[x for x in open('gpurun_out/%s.json'%f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1]); k=j['kernels']
    print('%-20s ms %.2f  dgrad %.2f fwd %.2f wgrad %.2f  grad_check %s allreduce %s' % (f, j['ms_per_step'], k['szn_conv_dgrad']['ms_per_step'], k['szn_conv_fwd']['ms_per_step'], k['szn_conv_wgrad']['ms_per_step'], (j.get('grad_check') or {}).get('worst_rel'), j.get('allreduce')))
PY
