#!/bin/bash
# usage: tools/scratch/variants.sh <variant dirs under build_variants/...>: quick device-resident bench of each libcrv_b200.so variant
cp cloudrendering_vulkan_b200/libcrv_b200.so /tmp/base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so; else cp build_variants/$v/libcrv_b200.so cloudrendering_vulkan_b200/libcrv_b200.so; fi
  for m in strict driver; do
    python bench.py --quick --no-e2e --no-cpu-baseline --math $m 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', '$m', round(d['value']), 'Mpath/s', round(d['ms_per_step'],3), 'ms; tracker', round(d['roofline']['kernel_ms_per_step'],3))"
  done
done
cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so
