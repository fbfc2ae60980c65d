#!/bin/bash
# One GPU call: the round-2 evidence that ends up (summarised) under profiles/.  Run as
#   gpurun --timeout 2400 -- 'bash tools/r02_capture.sh'
# Everything is written to gpurun_out/ (merged back); tools/summarize_profiles.py turns it into profiles/r02_*.
set -u
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/r02_clocks.csv &
SMI=$!
# (1) launch list of the timed steps of bench.py (cudaProfilerStart/Stop around them)
QCNN_PROFILE_RANGE=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-config4 --no-strict --no-cpu-baseline > $O/r02_bench_under_ncu.json 2> $O/r02_bench_under_ncu.err
# (2) the five decode-at-use GEMM launches of one forward pass at batch 256, full sections + source
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:pq_gemm_tc -c 5 -f -o $O/r02_pq_gemm_b256 \
    python tools/profile_step.py --batch 256 --iters 3 > $O/r02_prof_b256.log 2>&1
# (3) every kernel of the step: time, DRAM bytes, occupancy (metrics only: a --set full report of all launches is > 40 MB)
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none --profile-from-start off --csv --log-file $O/r02_step_b256_metrics.csv \
    python tools/profile_step.py --batch 256 --iters 3 > $O/r02_prof_b256_all.log 2>&1
# (4) batch 1: launch list and the persistent FC kernel, full sections + source
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $O/r02_launches_b1.csv \
    python tools/profile_step.py --batch 1 --iters 4 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fc_chain -c 1 -f -o $O/r02_fc_chain_b1 \
    python tools/profile_step.py --batch 1 --iters 4 > $O/r02_prof_b1.log 2>&1
kill $SMI
ls -la $O/*.ncu-rep; du -sh $O
