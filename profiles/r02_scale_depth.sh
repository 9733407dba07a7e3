#!/bin/bash
# Look-ahead depth of the two-phase scan across 8 GPUs (quick runs: timed step + config5 only).
mkdir -p gpurun_out
O=gpurun_out
for D in 2 1 3; do
  SB_CYCLIC_DEPTH=$D timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29800 + D)) bench.py --gpus 8 --steps 20 --warmup 5 \
     --no-e2e --no-replicas --no-cpu-baseline 2> $O/r02_scale_n8_D$D.err | tail -1 > $O/r02_scale_n8_D$D.json
  python - $O/r02_scale_n8_D$D.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read()); c = d["config5"]
print(sys.argv[1], "value", d["value"], "legs", {k: v["ms"] for k, v in d["roofline"]["per_leg"].items()}, "verified", d["verified"],
      "config5", c["ms_per_step"], {k: v["ms"] for k, v in c["per_leg"].items()}, "speedup", c.get("speedup_vs_1gpu"))
PY
done
