# ncu --set full captures (one launch each) of kernels that sit far from their roofline; reports come back in gpurun_out/
# (keep them small: the directory is capped at 64 MiB).
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-fused-head"
for K in ${KERNELS:-conv1_1_fwd_kernel conv1_1_wgrad_kernel pool_bwd_kernel}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_r02_$K $B > gpurun_out/ncu_$K.log 2>&1; echo "$K exit=$?"
done
# tensor-core conv launches of the 4th step (45 per step): index 0 = conv1_2 fwd, 4 = conv3_2 fwd, 43 / 44 = conv1_2 wgrad / dgrad
for IDX in ${UMMA_IDX:-0}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s $((135 + IDX)) -c 1 -f -o gpurun_out/prof_r02_umma_$IDX $B > gpurun_out/ncu_umma_$IDX.log 2>&1; echo "umma $IDX exit=$?"
done
du -sh gpurun_out
