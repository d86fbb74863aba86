#!/bin/bash
# The part of tools/gpu_profile.sh that bench.py's roofline depends on: the ncu metrics pass of the current sources
# (-> gpurun_out/<tag>_ncu_live.json, to be copied to profiles/ncu_live.json) followed by the default bench line.
# usage: tools/gpu_live.sh <tag> [extra bench args]
set -u
TAG=${1:-r2}; shift || true
OUT=gpurun_out; mkdir -p $OUT
SHA=$(python bench.py --print-source-sha); echo "$SHA" > $OUT/${TAG}_source_sha.txt
CFG=c2; prev=""; for a in "$@"; do [ "$prev" = "--config" ] && CFG=$a; prev=$a; done
METRICS=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed.avg.per_cycle_elapsed,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
timeout 1200 ncu --metrics $METRICS --clock-control none -k regex:"pt_wavefront|sv_fused|fold_kernel|group_|pt_classify" --csv \
    --log-file $OUT/${TAG}_live.csv python bench.py --quick --steps 1 --warmup 3 --no-e2e --no-cpu-baseline "$@" > $OUT/${TAG}_ncu_live.log 2>&1
python tools/ncu_summary.py live $OUT/${TAG}_live.csv $SHA $CFG > $OUT/${TAG}_ncu_live.json && cp $OUT/${TAG}_ncu_live.json profiles/ncu_live.json
timeout 900 python bench.py "$@" > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
cut -c1-300 $OUT/${TAG}_bench.json
