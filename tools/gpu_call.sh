# One short GPU session (run through gpurun); every step has its own timeout.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_full_size_gpu.py tests/test_model_gpu.py tests/test_trainer_gpu.py tests/test_train_step_gpu.py -q --timeout 150 > gpurun_out/pytest_final_subset.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/pytest_final_subset.log)"
timeout 60 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 gpurun_out/smoke.log)"
