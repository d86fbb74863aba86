#!/bin/bash
# Run on the GPU box (under gpurun): full default bench, ncu launch list of the same command, and one
# `ncu --set full` capture of the two hot kernels.  Outputs land in gpurun_out/<tag>_*.
# usage: tools/gpu_profile.sh <tag> [extra bench args]
set -u
TAG=${1:-r1}; shift || true
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > $OUT/${TAG}_clocks.csv &
SMI=$!
timeout 900 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
kill $SMI
cat $OUT/${TAG}_bench.json
timeout 900 python bench.py --impl reference "$@" > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json
# launch list of the same command (cold-cache, serialised: compare shares)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py "$@" > $OUT/${TAG}_bench_under_ncu.log 2>&1
# full capture of the hot kernels (short command: 16 spp)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pt_wavefront -s 3 -c 1 -f -o $OUT/${TAG}_pt \
    python bench.py --spp 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_pt.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sv_ -s 8 -c 2 -f -o $OUT/${TAG}_sv \
    python bench.py --spp 1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_sv.log 2>&1
# DRAM bytes of ONE tracker launch of the default bench step (256 frames per launch): the `traffic` of bench.py's roofline
timeout 1200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:pt_wavefront -s 2 -c 1 --csv \
    --log-file $OUT/${TAG}_pt_traffic.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_traffic.log 2>&1
ls -la $OUT
