#!/bin/bash
# Scaling record (run under `gpurun --gpus 8`): BASELINE configs[1] (weak) and configs[4] = 2^33 int32 (strong) at N = 1, 2, 4, 8.
# QUICK=1: skip the e2e and scan_then_carry extras (about 3 minutes of box time instead of 8).
mkdir -p gpurun_out
run() { # N workload extra...
  N=$1; W=$2; shift 2
  if [ "$N" = 1 ]; then python bench.py --gpus 1 --workload $W --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --workload $W --steps 10 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1
  fi
}
for N in 1 2 4 8; do run $N config2 --no-e2e; done > gpurun_out/r01_scale_config2.jsonl
for N in 1 2 4 8; do run $N config5 --no-e2e; done > gpurun_out/r01_scale_config5.jsonl
if [ -z "$QUICK" ]; then
  run 8 config5 --no-e2e --scan-strategy scan_then_carry > gpurun_out/r01_scale_config5_carry_n8.jsonl
  run 8 config2 > gpurun_out/r01_scale_config2_e2e_n8.jsonl
fi
