#!/bin/bash
# Scaling record (run under `gpurun --gpus 8`): the driver's own sequence — bench.py at N = 1, 2, 4, 8 with default flags
# (BASELINE configs[1] weak scaling + config5 / replicas sub-records + sharded e2e) — plus the 8-GPU tests.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r02_scale_topo.txt 2>&1
run() { # N tag extra...
  N=$1; TAG=$2; shift 2
  if [ "$N" = 1 ]; then timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" 2> $O/r02_scale_${TAG}.err | tail -1 > $O/r02_scale_${TAG}.json
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700 + N)) bench.py --gpus $N --steps 20 --warmup 5 "$@" 2> $O/r02_scale_${TAG}.err | tail -1 > $O/r02_scale_${TAG}.json
  fi
  python - "$O/r02_scale_${TAG}.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read())
    c5 = d.get("config5") or {}
    print(sys.argv[1], "N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "verified", d["verified"], "strategy", d["config"].get("scan_strategy"),
          "e2e", (d.get("e2e") or {}).get("value"), "config5 ms", c5.get("ms_per_step"), "speedup", c5.get("speedup_vs_1gpu"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
for N in 1 2 4 8; do run $N n$N --no-cpu-baseline; done
if [ -z "$QUICK" ]; then
  run 8 n8_rts --scan-strategy reduce_then_scan --no-e2e --no-replicas --no-cpu-baseline
  timeout 900 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py -x -q -m gpu -k "comm_scan_cyclic or foreach or binaries" > $O/r02_scale_pytest.txt 2>&1
  tail -3 $O/r02_scale_pytest.txt
fi
du -sh $O
