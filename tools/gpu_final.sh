#!/bin/bash
# Validation + evidence run of the final build of round 2 (one GPU).   gpurun --timeout 880 -- 'bash tools/gpu_final.sh'
mkdir -p gpurun_out
O=gpurun_out
OLD="SZN_POOL_CODE=0 SZN_CONV1_1_WGRAD_V2=0 SZN_COLSUM_GLOBAL=1"
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
sha256sum zeroshotsemanticsegmentation_b200/libszn.so | cut -c1-16 > $O/r02_so_hash.txt
python -c "from zeroshotsemanticsegmentation_b200 import _lib; print('build id (source hash, include/szn_build.h):', _lib.build_id())" >> $O/r02_so_hash.txt 2>&1
# 1. the kernels touched last, both forms of each, fail-fast
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k "conv1_1 or pool or dgrad" --timeout 200 > $O/next_kernel_tests.log 2>&1; echo "kernel tests exit=$? :: $(tail -1 $O/next_kernel_tests.log)"
SZN_COLSUM_GLOBAL=1 timeout 200 python -m pytest tests/test_kernels_gpu.py -q -k "dgrad" --timeout 200 > $O/next_kernel_tests_colsum_global.log 2>&1; echo "dgrad tests (global column sums) exit=$? :: $(tail -1 $O/next_kernel_tests_colsum_global.log)"
# 2. everything
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -s > $O/r02_pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -1 $O/r02_pytest_gpu.log)"
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 $O/r02_smoke.log)"
# 3. bench: the default; the column sums alone switched back; all three switched back
timeout 400 python bench.py > $O/r02_bench_config1_tf32.json 2> $O/bench.err; echo "cfg1 exit=$?"
SZN_COLSUM_GLOBAL=1 timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/r02_bench_config1_tf32_colsum_global.json 2>> $O/bench.err; echo "cfg1 colsum-global exit=$?"
env $OLD timeout 200 python bench.py --no-cpu-baseline --no-e2e > $O/r02_bench_config1_tf32_previous_kernels.json 2>> $O/bench.err; echo "cfg1 old exit=$?"
python - <<'P'
import json
for f in ("r02_bench_config1_tf32.json", "r02_bench_config1_tf32_colsum_global.json", "r02_bench_config1_tf32_previous_kernels.json"):
    try:
        d = json.loads(open("gpurun_out/" + f).read().strip().splitlines()[-1])
        k = d["kernels"]
        print(f, "ms/step %.3f" % d["ms_per_step"], "clk", d["clocks"]["sm_mhz"], {n: round(v["ms_per_step"], 3) for n, v in k.items() if "pool" in n or "conv1_1" in n or "conv_" in n})
    except Exception as e:
        print(f, "ERR", e)
P
# 4. profiles: launch lists, DRAM traffic of the conv family, --set full of the new kernels
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches_fused.log 2>&1
python tools/launch_summary.py $O/r02_launches_fused.csv 1 > $O/r02_launches_fused_summary.txt 2>&1; head -12 $O/r02_launches_fused_summary.txt
for K in pool_fwd_code_kernel pool_bwd_code_kernel conv1_1_wgrad_v2_kernel conv1_1_tc_kernel; do
  timeout 200 ncu --set full --clock-control none -k regex:$K -s 3 -c 1 -f -o $O/prof_r02_$K $B > $O/ncu_$K.log 2>&1
  python tools/ncu_metrics.py $O/prof_r02_$K.ncu-rep > $O/r02_ncu_full_${K}_summary.txt 2>&1; rm -f $O/prof_r02_$K.ncu-rep
  grep -E "gpu__time_duration|dram__bytes_read.sum |dram__bytes_write.sum " $O/r02_ncu_full_${K}_summary.txt | tr -s ' ' | tr '\n' ';'; echo " <- $K"
done
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:umma_conv_kernel -s 135 -c 45 --csv --log-file $O/r02_umma_traffic.csv $B > $O/ncu_traffic.log 2>&1
python tools/ncu_traffic.py $O/r02_umma_traffic.csv 1 tf32 > $O/r02_umma_traffic.json; tail -4 $O/r02_umma_traffic.json
# conv1_2 data gradient (the 45th tensor-core launch of a step) with the CTA-local column sums
timeout 200 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s $((135 + 44)) -c 1 -f -o $O/prof_r02_umma_44 $B > $O/ncu_umma_44.log 2>&1
python tools/ncu_metrics.py $O/prof_r02_umma_44.ncu-rep > $O/r02_ncu_full_umma_44_summary.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fused-head > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches.csv > $O/r02_launches_summary.txt 2>&1
# 5. probes: A operand from tensor memory; which cuDNN kernels win on the narrow layers
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_mma tools/probe_mma.cu > /dev/null 2>&1 && PROBE_TS=1 timeout 60 ./tools/probe_mma > $O/r02_probe_mma_ts.txt 2>&1; tail -11 $O/r02_probe_mma_ts.txt
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_cudnn_narrow_launches.csv python tools/library_layers.py tf32 8 3 conv1_2,conv2_1 > $O/ncu_cudnn.log 2>&1
python - <<'P'
import csv
rows = list(csv.DictReader(l for l in open("gpurun_out/r02_cudnn_narrow_launches.csv") if not l.startswith("==")))
names = [(r["Kernel Name"][:110], float(r["Metric Value"]) / 1e6) for r in rows]
# a winner is launched 3 (warm-up) + 3 (timed) times in a row at the end of its pass
out, i = [], 0
while i < len(names):
    j = i
    while j + 1 < len(names) and names[j + 1][0] == names[i][0]:
        j += 1
    if j - i + 1 >= 5:
        out.append("%2d x %-110s %.3f ms" % (j - i + 1, names[i][0], sum(t for _, t in names[i:j + 1]) / (j - i + 1)))
    i = j + 1
open("gpurun_out/r02_cudnn_narrow_kernels.txt", "w").write("\n".join(out) + "\n")
print("\n".join(out))
P
# 6. the other configs from the same build (no CPU arm)
timeout 200 python bench.py --precision fp32 --no-cpu-baseline --steps 20 > $O/r02_bench_config1_fp32grade.json 2>> $O/bench.err; echo "cfg1 fp32 exit=$?"
timeout 200 python bench.py --config 2 --steps 20 --no-cpu-baseline > $O/r02_bench_config2_bf16_B32.json 2>> $O/bench.err; echo "cfg2 exit=$?"
timeout 200 python bench.py --config 3 --steps 20 --no-cpu-baseline > $O/r02_bench_config3_zeroshot.json 2>> $O/bench.err; echo "cfg3 exit=$?"
timeout 200 python bench.py --config 4 --steps 20 --no-cpu-baseline > $O/r02_bench_config4_D1024_C256.json 2>> $O/bench.err; echo "cfg4 exit=$?"
timeout 200 python bench.py --config 0 --no-cpu-baseline > $O/r02_bench_config0.json 2>> $O/bench.err; echo "cfg0 exit=$?"
rm -f $O/ncu_*.log $O/r02_cudnn_narrow_launches.csv
du -sh $O
