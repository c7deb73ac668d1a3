# One short GPU session (run through gpurun); every step has its own timeout.
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_edge_cases_gpu.py -q --timeout 100 > gpurun_out/pytest_edge.log 2>&1; echo "edge exit=$? :: $(tail -1 gpurun_out/pytest_edge.log)"
timeout 120 python tools/library_bar.py --upscore grouped --dtype tf32 > gpurun_out/library_tf32.log 2>&1; echo "lib tf32 exit=$? :: $(tail -c 400 gpurun_out/library_tf32.log)"
timeout 80 python tools/library_bar.py --upscore grouped --dtype bf16 > gpurun_out/library_bf16.log 2>&1; echo "lib bf16 exit=$? :: $(tail -c 400 gpurun_out/library_bf16.log)"
