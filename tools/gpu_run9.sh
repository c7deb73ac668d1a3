mkdir -p gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
A="bench.py --gpus 2 --steps 30 --warmup 3 --no-e2e --no-grad-check"
timeout 300 python bench.py --no-cpu-baseline --steps 30 --no-e2e > gpurun_out/sc_n1.json 2> gpurun_out/sc.err
SZN_STATIC_TILES=1 SZN_DDP_WGRAD_WAVES=1 timeout 300 $R 29511 $A > gpurun_out/sc_static_w1.json 2>> gpurun_out/sc.err
SZN_DDP_WGRAD_WAVES=1 timeout 300 $R 29512 $A > gpurun_out/sc_dyn_w1.json 2>> gpurun_out/sc.err
SZN_DDP_WGRAD_WAVES=2 timeout 300 $R 29513 $A > gpurun_out/sc_dyn_w2.json 2>> gpurun_out/sc.err
SZN_DDP_WGRAD_WAVES=3 timeout 300 $R 29514 $A > gpurun_out/sc_dyn_w3.json 2>> gpurun_out/sc.err
SZN_STATIC_TILES=1 SZN_DDP_WGRAD_WAVES=1 timeout 300 $R 29515 $A > gpurun_out/sc_static_w1_again.json 2>> gpurun_out/sc.err
python - <<'PY'
import json,glob
for f in ['sc_n1','sc_static_w1','sc_dyn_w1','sc_dyn_w2','sc_dyn_w3','sc_static_w1_again']:
    l=[x for x in open('gpurun_out/%s.json'%f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1]); k=j['kernels']
    print('%-20s ms %.2f  dgrad %.2f fwd %.2f wgrad %.2f' % (f, j['ms_per_step'], k['szn_conv_dgrad']['ms_per_step'], k['szn_conv_fwd']['ms_per_step'], k['szn_conv_wgrad']['ms_per_step']))
PY
