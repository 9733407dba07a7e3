#!/bin/bash
# Round 2, GPU call 11 (2 GPUs): publish-ahead look-back warp of the two-phase scan — parity + bench N=2.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py -x -q -m gpu > $O/r02_call11_pytest.txt 2>&1
tail -3 $O/r02_call11_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 2 --steps 20 --warmup 5 \
     --no-e2e --no-replicas 2> $O/r02_bench_n2_pa.err | tail -1 > $O/r02_bench_n2_pa.json
python - $O/r02_bench_n2_pa.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read()); c = d["config5"]
print("N=2 value", d["value"], "legs", {k: v["ms"] for k, v in d["roofline"]["per_leg"].items()}, "verified", d["verified"],
      "config5", c["ms_per_step"], {k: v["ms"] for k, v in c["per_leg"].items()}, "speedup", c.get("speedup_vs_1gpu"))
PY
timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_cyclic > $O/r02_ops_v7.jsonl 2>> $O/r02_call11.err
python profiles/fmt_jsonl.py $O/r02_ops_v7.jsonl
