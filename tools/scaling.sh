#!/bin/bash
# 1/2/4/8-GPU run of bench.py on one box (under `gpurun --gpus 8`); one JSON line per N in gpurun_out/<tag>_scale.jsonl.
# bench.py's default is strong scaling (the config's spp split over the GPUs); the weak figure travels in `scaling_other`.
TAG=${1:-r2}; shift || true
OUT=gpurun_out; mkdir -p $OUT; : > $OUT/${TAG}_scale.jsonl
for N in ${CRV_SCALE_NS:-1 2 4 8}; do
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
for r in rows:
    e=r.get("e2e") or {}; o=r.get("scaling_other") or {}
    print(r["n_gpus"], r["scaling"], round(r["value"]), "Mpath-samples/s", "x%.2f" % (r["value"]/base), "ms/step", round(r["ms_per_step"],3),
          "tracker", round(r["roofline"]["kernel_ms_per_step"],3), "combine", {k:(round(v,4) if isinstance(v,float) else v) for k,v in r["combine"].items() if k!="note"},
          "e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),2), "other", o.get("scaling"), round(o.get("value",0)))
PY
