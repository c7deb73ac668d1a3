#!/bin/bash
# 8-GPU session (gpurun --gpus 8): config 1 scaling points + A/B of the knobs that matter under NCCL overlap, then
# BASELINE configs[3] and [4] at their quoted 8-GPU size.  JSON lines land in gpurun_out/ (copied to profiles/r02_*).
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # name nproc port env... -- bench args
  local name=$1 n=$2 port=$3; shift 3
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 400 $R --nproc-per-node $n --master-port $port bench.py --gpus $n "$@" > $O/$name.json 2>> $O/scale8.err
  echo "$name exit=$?"
}
timeout 300 python bench.py --no-cpu-baseline --steps 30 > $O/r02_scale_cfg1_n1.json 2> $O/scale8.err; echo "n1 exit=$?"
run r02_scale_cfg1_n8 8 29601 X=1 -- --steps 30 --warmup 3
run r02_scale_cfg1_n4 4 29602 X=1 -- --steps 30 --warmup 3
run r02_scale_cfg1_n2 2 29603 X=1 -- --steps 30 --warmup 3
run ab_n8_static 8 29604 SZN_STATIC_TILES=1 -- --steps 20 --warmup 3 --no-e2e --no-grad-check
run ab_n8_waves3 8 29605 SZN_DDP_WGRAD_WAVES=3 -- --steps 20 --warmup 3 --no-e2e --no-grad-check
run ab_n8_maxctas8 8 29606 NCCL_MAX_CTAS=8 -- --steps 20 --warmup 3 --no-e2e --no-grad-check
run r02_bench_config3_zeroshot_8gpu 8 29607 X=1 -- --config 3 --steps 20 --warmup 3
run r02_bench_config4_D1024_C256_8gpu 8 29608 X=1 -- --config 4 --steps 20 --warmup 3
run r02_bench_config2_bf16_B32_8gpu 8 29609 X=1 -- --config 2 --steps 20 --warmup 3
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02_scale_cfg1_n*.json'))+sorted(glob.glob('gpurun_out/ab_n8_*.json'))+sorted(glob.glob('gpurun_out/r02_bench_config*_8gpu.json')):
    l=[x for x in open(f) if x.startswith('{')]
    if not l: print(f,'NO JSON'); continue
    j=json.loads(l[-1]); k=j['kernels']
    print('%-48s n=%d value %7.1f ms %.2f e2e %s grad_check %s | dgrad %.2f fwd %.2f wgrad %.2f' % (f.split('/')[-1], j['n_gpus'], j['value'], j['ms_per_step'], ('%.1f' % j['e2e']['value']) if 'e2e' in j else '-', (j.get('grad_check') or {}).get('worst_rel'), k['szn_conv_dgrad']['ms_per_step'], k['szn_conv_fwd']['ms_per_step'], k['szn_conv_wgrad']['ms_per_step']))
PY
tail -c 600 $O/scale8.err
