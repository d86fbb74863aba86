#!/bin/bash
# compute-sanitizer passes over the hot kernels (run on the GPU box): memcheck + racecheck on the smoke invocation, memcheck on
# the tests that cover what round 2 added (job orders, sorted hit list, shadow-volume build modes incl. the row kernel, forced
# coarse-table shapes).  The group combines are left out: their stream-ordered waits between members need launches that
# stay asynchronous, which the sanitizer does not guarantee.  Summary lines -> gpurun_out/<tag>_sanitizer.txt
TAG=${1:-r2}; OUT=gpurun_out; mkdir -p $OUT
S=$OUT/${TAG}_sanitizer.txt; : > $S
run() {  # label, tool, command...
  local label=$1 tool=$2; shift 2
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 "$@" > $OUT/${TAG}_san_tmp.log 2>&1
  echo "$tool  $label  (rc $?) -> $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/${TAG}_san_tmp.log | tail -1)   [$(grep -E '[0-9]+ passed|[0-9]+ failed' $OUT/${TAG}_san_tmp.log | tail -1)]" >> $S
  grep -E "Invalid|hazard|Error:" $OUT/${TAG}_san_tmp.log | head -5 >> $S
}
run "python __graft_entry__.py smoke" memcheck python __graft_entry__.py smoke
run "python __graft_entry__.py smoke" racecheck python __graft_entry__.py smoke
run "pytest tests/test_gpu_parity.py -k 'job_orders or shadow_volume_modes or batch_equals or empty_cloud'" memcheck \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "job_orders or shadow_volume_modes or batch_equals or empty_cloud"
run "pytest tests/test_production_shapes.py -k 'coarse_table_shapes or rebind'" memcheck \
    python -m pytest tests/test_production_shapes.py -x -q -m gpu -k "coarse_table_shapes or rebind"
run "pytest tests/test_gpu_parity.py -k 'shadow_volume_modes' (row kernel: shuffles)" racecheck \
    python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "shadow_volume_modes"
rm -f $OUT/${TAG}_san_tmp.log
cat $S
