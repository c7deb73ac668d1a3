#!/bin/bash
# One GPU session: kernel-level parity tests in isolated processes (a trap in one group cannot poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 300 python -m pytest "$@" -q --timeout 120 -s > gpurun_out/$name.log 2>&1
  echo "$name exit=$? :: $(tail -1 gpurun_out/$name.log)"
  grep -E "relerr|rel err|^FAILED|^ERROR" gpurun_out/$name.log | head -40
}
run k_fwd   tests/test_kernels_gpu.py -k "test_conv_fwd"
run k_dgrad tests/test_kernels_gpu.py -k "test_conv_dgrad"
run k_wgrad tests/test_kernels_gpu.py -k "test_conv_wgrad"
run k_misc  tests/test_kernels_gpu.py -k "conv1_1 or pool or bias_grad or upsample"
run head    tests/test_head_gpu.py
run model   tests/test_model_gpu.py
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; echo "bench exit=$? :: $(tail -c 3000 gpurun_out/bench_quick.log)"
