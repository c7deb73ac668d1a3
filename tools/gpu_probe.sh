#!/bin/bash
# Round-2 first GPU call: the questions DESIGN §6 "Next" leaves open, answered by measurement (≈1 GPU-minute).
#   gpurun --timeout 300 -- 'bash tools/gpu_probe.sh'
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_mma tools/probe_mma.cu > gpurun_out/probe_build.log 2>&1 \
  && timeout 120 ./tools/probe_mma > gpurun_out/probe_mma.txt 2>&1
echo "probe exit=$?"; head -70 gpurun_out/probe_mma.txt
timeout 120 python tools/bench_layers.py tf32 8 5 > gpurun_out/layers_tf32.txt 2>&1; tail -25 gpurun_out/layers_tf32.txt
