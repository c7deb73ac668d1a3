# One short GPU session (run through gpurun); every step has its own timeout.
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/pytest_gpu.log)"
timeout 200 python bench.py --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_steps40.log 2>&1; echo "bench40 exit=$? :: $(tail -c 300 gpurun_out/bench_steps40.log)"
timeout 200 python bench.py --config 4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4.log 2>&1; echo "cfg4 exit=$? :: $(tail -c 300 gpurun_out/bench_cfg4.log)"
