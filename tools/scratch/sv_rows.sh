#!/bin/bash
cp cloudrendering_vulkan_b200/libcrv_b200.so /tmp/base.so
for v in base "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so; else cp build_variants/$v/libcrv_b200.so cloudrendering_vulkan_b200/libcrv_b200.so; fi
  python tools/sv_bench.py --layout brick --mode fused_rows --cases 512:512,1024:1024,500:128 | python -c "import sys,json; print('$v rows', *[(d['n_sv'], round(d['ms_median'],4)) for d in map(json.loads, sys.stdin)])"
done
cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so
python tools/sv_bench.py --layout brick --mode fused --cases 512:512,1024:1024,500:128 | python -c "import sys,json; print('fused (cta) ', *[(d['n_sv'], round(d['ms_median'],4)) for d in map(json.loads, sys.stdin)])"
