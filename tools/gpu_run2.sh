mkdir -p gpurun_out
timeout 60 ./tools/probe_mma > gpurun_out/probe_mma_r02.txt 2>&1; echo "probe exit=$?"; cat gpurun_out/probe_mma_r02.txt | head -80
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_head_gpu.py tests/test_edge_cases_gpu.py -q -x --timeout 120 2>&1 | tail -5
timeout 120 python tools/bench_layers.py tf32 8 5 > gpurun_out/layers_tf32_r02a.txt 2>&1; tail -20 gpurun_out/layers_tf32_r02a.txt
timeout 120 python tools/bench_layers.py bf16 8 5 > gpurun_out/layers_bf16_r02a.txt 2>&1; tail -20 gpurun_out/layers_bf16_r02a.txt
bash tools/gpu_quick.sh
