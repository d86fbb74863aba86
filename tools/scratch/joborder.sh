#!/bin/bash
# A/B of the tracker's job order on one GPU: the 32-frame launch of an 8-GPU strong-scaling member and the full 256-frame launch
for spp in 32 256; do for o in 0 1; do for m in strict; do
  python bench.py --quick --no-e2e --no-cpu-baseline --spp $spp --job-order $o --math $m --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('spp $spp order $o $m', round(d['value']), 'Mpath/s', round(d['ms_per_step'],4), 'ms; tracker', round(d['roofline']['kernel_ms_per_step'],4))"
done; done; done
