#!/bin/bash
# 1/2/4/8-GPU weak-scaling run of bench.py on one box (under `gpurun --gpus 8`); one JSON line per N in gpurun_out/<tag>_scale.jsonl
TAG=${1:-r1}; shift || true
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_scale.jsonl
for N in 1 2 4 8; do
  if [ $N = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline "$@" >> $OUT/${TAG}_scale.jsonl 2> $OUT/${TAG}_scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
      bench.py --gpus $N --steps 5 --warmup 3 "$@" >> $OUT/${TAG}_scale.jsonl 2> $OUT/${TAG}_scale_n$N.err
  fi
done
python - <<PY
import json
rows=[json.loads(l) for l in open("$OUT/${TAG}_scale.jsonl") if l.startswith("{")]
base=rows[0]["value"]
for r in rows: print(r["n_gpus"], round(r["value"]), "Mpath-samples/s", "x%.2f" % (r["value"]/base), "e2e", round(r["e2e"]["value"]), "ms/step", round(r["ms_per_step"],2))
PY
