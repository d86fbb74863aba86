#!/bin/bash
# the CRV_PT_TIMELINE measurement build (build_variants/tl) on the 32-frame launch, both job orders
cp cloudrendering_vulkan_b200/libcrv_b200.so /tmp/base.so
cp build_variants/tl/libcrv_b200.so cloudrendering_vulkan_b200/libcrv_b200.so
for o in 0 1; do echo "job order $o"; python bench.py --quick --no-e2e --no-cpu-baseline --spp 32 --job-order $o --steps 3 2>&1 | grep "pt timeline\|loop iterations" | tail -4; done
cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so
