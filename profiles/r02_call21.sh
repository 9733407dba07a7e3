#!/bin/bash
# Round 2, GPU call 21 (2 GPUs): regression check of the multi-GPU paths on the final build — the 2-GPU tests of the C ABI
# (sb_comm_*, sb_scan_cyclic, replicas, binaries with --num-gpus=2) and bench.py at N = 2 as the driver launches it;
# kmeans after the fma / early-exit changes.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py tests/test_single_kernel_widened.py -q -m gpu --maxfail=10 ) > $O/r02_call21_pytest.txt 2>&1
tail -4 $O/r02_call21_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline 2> $O/r02_bench_n2_final.err | tail -1 > $O/r02_bench_n2_final.json
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_n2_final.json').read())
c5 = d.get("config5") or {}
print("N", d["n_gpus"], "value", d["value"], "ms", d["ms_per_step"], "verified", d["verified"], d["verified_legs"], "e2e", (d.get("e2e") or {}).get("value"),
      "config5 ms", c5.get("ms_per_step"), "speedup", c5.get("speedup_vs_1gpu"))
print(d["roofline"]["per_leg"]); print(d["replicas"]["config3"]["per_leg"])
PY
B=sycl-bench_b200/bin
rm -f $O/r02_kmeans_v4.csv
timeout 600 $B/kmeans --size=268435456 --local=256 --num-runs=10 --warmup-run --no-verification --output=$O/r02_kmeans_v4.csv > /dev/null 2>> $O/r02_call21.err
grep -v "^#" $O/r02_kmeans_v4.csv | cut -d, -f1,8 
