#!/bin/bash
# the CRV_PT_EXACT_STATS measurement build (build_variants/ex): what becomes of the tentative collisions the coarse table passes on
cp cloudrendering_vulkan_b200/libcrv_b200.so /tmp/base.so
cp build_variants/ex/libcrv_b200.so cloudrendering_vulkan_b200/libcrv_b200.so
for cfg in "c2 --spp 32" "c4 --spp 16" "c3 --spp 8"; do echo "$cfg"; python bench.py --quick --no-e2e --no-cpu-baseline --config $cfg --steps 1 --warmup 3 2>&1 | grep "exact-stage census" | tail -1; done
cp /tmp/base.so cloudrendering_vulkan_b200/libcrv_b200.so
