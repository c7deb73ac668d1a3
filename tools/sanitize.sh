#!/bin/bash
# compute-sanitizer over the kernel parity tests (memcheck / racecheck / synccheck for the hand-rolled
# mbarrier / TMEM / TMA pipelines).  Run on the GPU box:  gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# The sanitizer slows every launch ~10-100x: a representative subset (every C-ABI kernel at least once, all three
# precisions for the tensor-core conv) keeps it to a few minutes.  Logs land in gpurun_out/ (summaries go to profiles/).
mkdir -p gpurun_out
SEL='test_conv_fwd and case0 or test_conv_dgrad and case1 or test_conv_wgrad and case2 or test_conv_dgrad_col2im or test_conv1_1 or test_pool and hw0 or test_bias_grad or test_upsample_and_small_deconv and HW0 or test_conv_fwd_fp32_out'
for TOOL in memcheck racecheck synccheck; do
  timeout 700 compute-sanitizer --tool $TOOL --error-exitcode 99 --print-limit 20 \
    python -m pytest tests/test_kernels_gpu.py tests/test_head_gpu.py -q -x --timeout 600 -k "$SEL or test_head" -p no:cacheprovider \
    > gpurun_out/sanitizer_$TOOL.log 2>&1
  echo "$TOOL exit=$? :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/sanitizer_$TOOL.log | tail -3 | tr '\n' ' ')"
done
