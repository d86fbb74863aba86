#!/bin/bash
# Run on the GPU box (under gpurun): full default bench + reference arm, ncu launch list of the same command, one
# `ncu --set full` capture of the two hot kernels, and the metrics pass bench.py's roofline reads (profiles/ncu_live.json,
# stamped with the hash of the kernel sources).  Outputs land in gpurun_out/<tag>_*; tools/ncu_summary.py turns them into the
# tracked summaries under profiles/.
# usage: tools/gpu_profile.sh <tag> [extra bench args, e.g. --config c2 --math strict]
set -u
TAG=${1:-r2}; shift || true
OUT=gpurun_out
mkdir -p $OUT
SHA=$(python bench.py --print-source-sha)
echo "$SHA" > $OUT/${TAG}_source_sha.txt
CFG=c2
prev=""; for a in "$@"; do [ "$prev" = "--config" ] && CFG=$a; prev=$a; done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 200 > $OUT/${TAG}_clocks.csv &
SMI=$!
# metrics pass first: bench.py below then reports the issue / DRAM fractions of THIS build
METRICS=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 1200 ncu --metrics $METRICS --clock-control none -k regex:"pt_wavefront|sv_fused|fold_kernel|group_|pt_classify" --csv \
    --log-file $OUT/${TAG}_live.csv python bench.py --quick --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_live.log 2>&1
python tools/ncu_summary.py live $OUT/${TAG}_live.csv $SHA $CFG > $OUT/${TAG}_ncu_live.json && cp $OUT/${TAG}_ncu_live.json profiles/ncu_live.json
timeout 900 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
kill $SMI
cat $OUT/${TAG}_bench.json
timeout 900 python bench.py --impl reference "$@" > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err
cat $OUT/${TAG}_bench_reference.json
# launch list of the same command (cold-cache, serialised: compare shares)
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --quick --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_bench_under_ncu.log 2>&1
# full capture of the hot kernels
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:pt_wavefront -s 3 -c 1 -f -o $OUT/${TAG}_pt \
    python bench.py --quick --spp 32 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_pt.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:sv_ -s 8 -c 2 -f -o $OUT/${TAG}_sv \
    python bench.py --quick --spp 1 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_sv.log 2>&1
ls -la $OUT
