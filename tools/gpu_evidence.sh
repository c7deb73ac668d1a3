#!/bin/bash
# Round-2 evidence run (one GPU): everything profiles/ cites, from ONE build of libszn.so.
#   gpurun --timeout 3000 -- 'bash tools/gpu_evidence.sh'
# Part 1: tests + smoke.  Part 2: bench lines of all five BASELINE configs (+ the fp32-grade mode).  Part 3: ncu launch list,
# DRAM traffic of the conv family, --set full captures (summaries are extracted on the box, only two reports travel back).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
sha256sum zeroshotsemanticsegmentation_b200/libszn.so | cut -c1-16 > $O/r02_so_hash.txt
python -c "from zeroshotsemanticsegmentation_b200 import _lib; print('build id (source hash, include/szn_build.h):', _lib.build_id())" >> $O/r02_so_hash.txt 2>&1
if [ "${PART:-all}" = "all" ] || [ "$PART" = "1" ]; then
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -s > $O/r02_pytest_gpu.log 2>&1; echo "pytest exit=$? :: $(tail -1 $O/r02_pytest_gpu.log)"
timeout 120 python __graft_entry__.py smoke > $O/r02_smoke.log 2>&1; echo "smoke exit=$? :: $(tail -1 $O/r02_smoke.log)"
fi
if [ "${PART:-all}" = "all" ] || [ "$PART" = "2" ]; then
timeout 400 python bench.py > $O/r02_bench_config1_tf32.json 2> $O/bench.err; echo "cfg1 exit=$?"
timeout 300 python bench.py --precision fp32 > $O/r02_bench_config1_fp32grade.json 2>> $O/bench.err; echo "cfg1 fp32 exit=$?"
timeout 300 python bench.py --config 0 > $O/r02_bench_config0.json 2>> $O/bench.err; echo "cfg0 exit=$?"
timeout 300 python bench.py --config 2 --steps 20 > $O/r02_bench_config2_bf16_B32.json 2>> $O/bench.err; echo "cfg2 exit=$?"
timeout 300 python bench.py --config 3 --steps 20 > $O/r02_bench_config3_zeroshot.json 2>> $O/bench.err; echo "cfg3 exit=$?"
timeout 300 python bench.py --config 4 --steps 20 > $O/r02_bench_config4_D1024_C256.json 2>> $O/bench.err; echo "cfg4 exit=$?"
timeout 120 python tools/bench_layers.py tf32 8 5 > $O/r02_layers_tf32.txt 2>&1
timeout 120 python tools/bench_layers.py bf16 8 5 > $O/r02_layers_bf16.txt 2>&1
timeout 120 python tools/bench_layers.py fp32 8 5 > $O/r02_layers_fp32grade.txt 2>&1
timeout 300 python tools/library_layers.py tf32 8 5 > $O/r02_library_layers_tf32.txt 2>&1; echo "library layers exit=$?"
timeout 300 python tools/library_layers.py bf16 8 5 > $O/r02_library_layers_bf16.txt 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/probe_mma tools/probe_mma.cu > /dev/null 2>&1 && timeout 60 ./tools/probe_mma > $O/r02_probe_mma.txt 2>&1
fi
if [ "${PART:-all}" = "all" ] || [ "$PART" = "3" ]; then
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-fused-head > $O/ncu_launches.log 2>&1
python tools/launch_summary.py $O/r02_launches.csv > $O/r02_launches_summary.txt 2>&1; head -24 $O/r02_launches_summary.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_fused.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_launches_fused.log 2>&1
python tools/launch_summary.py $O/r02_launches_fused.csv 1 > $O/r02_launches_fused_summary.txt 2>&1; head -8 $O/r02_launches_fused_summary.txt
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:umma_conv_kernel -s 135 -c 45 --csv --log-file $O/r02_umma_traffic.csv $B > $O/ncu_traffic.log 2>&1
python tools/ncu_traffic.py $O/r02_umma_traffic.csv 1 tf32 > $O/r02_umma_traffic.json; tail -4 $O/r02_umma_traffic.json
for IDX in 0 4 35 36 43 44; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:umma_conv_kernel -s $((135 + IDX)) -c 1 -f -o $O/prof_r02_umma_$IDX $B > $O/ncu_umma_$IDX.log 2>&1
  python tools/ncu_metrics.py $O/prof_r02_umma_$IDX.ncu-rep > $O/r02_ncu_full_umma_${IDX}_summary.txt 2>&1
  [ $IDX != 0 ] && [ $IDX != 44 ] && rm -f $O/prof_r02_umma_$IDX.ncu-rep
done
for K in conv1_1_tc_kernel conv1_1_wgrad_v2_kernel pool_fwd_code_kernel pool_bwd_code_kernel upsample_fwd_kernel upsample_bwd_kernel embed_loss_fwd_kernel embed_loss_bwd_kernel embed_argmax_tc_kernel; do
  timeout 300 ncu --set full --clock-control none -k regex:$K -s 3 -c 1 -f -o $O/prof_r02_$K $B --no-fused-head > $O/ncu_$K.log 2>&1
  python tools/ncu_metrics.py $O/prof_r02_$K.ncu-rep > $O/r02_ncu_full_${K}_summary.txt 2>&1; rm -f $O/prof_r02_$K.ncu-rep
done
timeout 300 ncu --set full --clock-control none -k regex:fused_pixels_kernel -s 3 -c 1 -f -o $O/prof_r02_fused_pixels $B > $O/ncu_fused_pixels.log 2>&1
python tools/ncu_metrics.py $O/prof_r02_fused_pixels.ncu-rep > $O/r02_ncu_full_fused_pixels_summary.txt 2>&1; rm -f $O/prof_r02_fused_pixels.ncu-rep
fi
if [ "$PART" = "4" ]; then
timeout 900 python bench.py --impl reference --cpu-as-written --steps 1 --warmup 0 > $O/r02_cpu_reference_as_written.json 2> $O/cpu_as_written.err; echo "as-written exit=$?"; cat $O/r02_cpu_reference_as_written.json | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_cpu_reference.json 2>> $O/cpu_as_written.err; cat $O/r02_cpu_reference.json | cut -c1-300
fi
rm -f $O/ncu_*.log
du -sh $O
