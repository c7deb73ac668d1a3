mkdir -p gpurun_out
timeout 60 ./tools/probe_mma > gpurun_out/probe_mma_r02b.txt 2>&1; echo "probe exit=$?"; grep -A20 "cta_group::2" gpurun_out/probe_mma_r02b.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02_bench_cfg1_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "2gpu exit=$?"; tail -c 800 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r02_bench_cfg1_2gpu.json') if x.startswith('{')]
if l:
    j=json.loads(l[-1]); print('2gpu value %.1f ms %.2f e2e %.1f grad_check %s allreduce %s' % (j['value'], j['ms_per_step'], j['e2e']['value'], j.get('grad_check'), j.get('allreduce_bytes_per_step')))
PY
timeout 300 python -m pytest tests/test_ddp_gpu.py -q --timeout 200 -s 2>&1 | tail -4
