mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_model_gpu.py -q -x --timeout 200 -k "fp32_grade" -s 2>&1 | grep -E "rel|passed|failed|assert" | tail -25
timeout 900 python -m pytest tests/test_full_size_gpu.py -q --timeout 600 -k "forward_vs_oracle" -s 2>&1 | grep -E "full-size|agreement|passed|failed|Error|assert" | tail -40
SZN_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_full_size_gpu.py::test_model_full_size_forward_vs_oracle_and_batch_consistency > gpurun_out/pytest_gpu_r02b.log 2>&1; echo "pytest exit=$?"; tail -8 gpurun_out/pytest_gpu_r02b.log
timeout 120 python tools/bench_layers.py fp32 8 5 > gpurun_out/layers_fp32_r02b.txt 2>&1; tail -18 gpurun_out/layers_fp32_r02b.txt
timeout 120 python tools/bench_layers.py tf32 8 5 > gpurun_out/layers_tf32_r02b.txt 2>&1; tail -3 gpurun_out/layers_tf32_r02b.txt
