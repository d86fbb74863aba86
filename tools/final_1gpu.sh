#!/bin/bash
# final single-GPU pass of a round: profile + bench + reference arm + ncu, GPU suite, other configs, sanitizer
TAG=${1:-r2q}
tools/gpu_profile.sh $TAG > gpurun_out/${TAG}_profile.log 2>&1
( time python -m pytest tests -q -m gpu 2>&1 | tail -3 ) > gpurun_out/${TAG}_pytest.log 2>&1
for c in c1_hg c4 c5 c3; do timeout 900 python bench.py --config $c > gpurun_out/${TAG}_$c.json 2> gpurun_out/${TAG}_$c.err; done
tools/sanitize.sh $TAG > /dev/null 2>&1
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_sanitizer.txt; cut -c1-400 gpurun_out/${TAG}_bench.json
