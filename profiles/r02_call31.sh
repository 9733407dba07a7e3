#!/bin/bash
# Round 2, GPU call 31 (8 GPUs): bench.py at N = 8 exactly as the driver launches it, on the final build (two-phase scan kernel on
# 14 x 10 tiles).  One run only: the 1 / 2 / 4-GPU records of the final build are profiles/r02_bench_n1_final.json, r02_bench_n2_final.json.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r02_bench_n8_final.err | tail -1 > $O/r02_bench_n8_final.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n8_final.json').read())
c5 = d.get("config5") or {}
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "verified", d["verified"], d["verified_legs"], "e2e", (d.get("e2e") or {}).get("value"),
      "config5 ms", c5.get("ms_per_step"), "speedup", c5.get("speedup_vs_1gpu"), c5.get("per_leg"))
print(d["roofline"]["per_leg"]); print(d["replicas"]["config3"]["value"], d["replicas"]["config4"]["value"])
PY
