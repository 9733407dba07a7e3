#!/bin/bash
# Round 2, GPU call 30 (2 GPUs): two-phase scan kernel on 14 x 10 tiles: the cyclic / comm tests and bench.py at N = 2.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py -q -m gpu --maxfail=10 ) > $O/r02_call30_pytest.txt 2>&1
tail -4 $O/r02_call30_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r02_bench_n2_final.err | tail -1 > $O/r02_bench_n2_final.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n2_final.json').read())
c5 = d.get("config5") or {}
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "verified", d["verified"], d["verified_legs"], "e2e", (d.get("e2e") or {}).get("value"),
      "config5 ms", c5.get("ms_per_step"), "speedup", c5.get("speedup_vs_1gpu"), c5.get("per_leg"))
print(d["roofline"]["per_leg"])
PY
