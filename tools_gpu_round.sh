#!/bin/bash
# One GPU session: kernel-level parity tests in isolated processes (a trap in one group cannot poison the others).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, pytest args...
  local name=$1; shift
  timeout 300 python -m pytest "$@" -q -x --timeout 120 -s > gpurun_out/$name.log 2>&1
  echo "$name exit=$? :: $(tail -1 gpurun_out/$name.log)"
}
run k_fwd_tf32   tests/test_kernels_gpu.py -k "test_conv_fwd and tf32"
run k_fwd_bf16   tests/test_kernels_gpu.py -k "test_conv_fwd and bf16"
run k_dgrad_tf32 tests/test_kernels_gpu.py -k "test_conv_dgrad and tf32"
run k_dgrad_bf16 tests/test_kernels_gpu.py -k "test_conv_dgrad and bf16"
run k_wgrad_tf32 tests/test_kernels_gpu.py -k "test_conv_wgrad and tf32"
run k_wgrad_bf16 tests/test_kernels_gpu.py -k "test_conv_wgrad and bf16"
run k_misc       tests/test_kernels_gpu.py -k "conv1_1 or pool or bias_grad or upsample"
run head         tests/test_head_gpu.py
run model        tests/test_model_gpu.py
