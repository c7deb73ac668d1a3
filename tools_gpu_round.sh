#!/bin/bash
# One GPU session (run through gpurun): parity tests, smoke, default bench and the ncu launch list.  Every step has its
# own short timeout: a hung kernel must not eat the GPU budget.
#   gpurun --timeout 900 -- 'bash tools_gpu_round.sh'            the validated default path
#   gpurun --timeout 900 -- 'bash tools_gpu_round.sh experimental'  + the opt-in fused head (DESIGN §6 "Next", item 2)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
timeout 420 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 gpurun_out/smoke.log)"
timeout 400 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench exit=$? :: $(tail -c 400 gpurun_out/bench_default.log)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -20 gpurun_out/launches_summary.txt
if [ "$1" = "experimental" ]; then
  SZN_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_fused_head_gpu.py -q --timeout 100 -s > gpurun_out/pytest_fused_head.log 2>&1
  echo "fused-head tests exit=$? :: $(tail -1 gpurun_out/pytest_fused_head.log)"
  timeout 200 python bench.py --fused-head --no-cpu-baseline > gpurun_out/bench_fused_head.log 2>&1; echo "fused bench exit=$? :: $(tail -c 300 gpurun_out/bench_fused_head.log)"
  bash tools/gpu_probe.sh
fi
