# One short GPU session (run through gpurun); every step has its own timeout.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q --timeout 200 -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -1 gpurun_out/pytest_gpu.log)"
grep -a "full-size forward\|label agreement with" gpurun_out/pytest_gpu.log | grep -v print
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches_summary.txt 2>&1; head -8 gpurun_out/launches_summary.txt
