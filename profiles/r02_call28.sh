#!/bin/bash
# Round 2, GPU call 28 (1 GPU): per-type tile shapes of the plain scans: full GPU suite, op table, bench.py (driver's command).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call28_pytest_all.txt 2>&1
tail -4 $O/r02_call28_pytest_all.txt
: > $O/r02_ops_v22.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v22.jsonl 2>> $O/r02_call28.err; }
run --ops scan
SB_SCAN_CYCLIC=0 run --ops scan --dtypes i32,i64 --log2n 30
python profiles/fmt_jsonl.py $O/r02_ops_v22.jsonl
( time timeout 900 python bench.py ) > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[0])
print({k: l[k] for k in ("value", "ms_per_step", "verified", "gpu_launches")}, l["roofline"]["frac"], l["roofline"]["per_leg"])
print(l["replicas"]["config3"]["per_leg"]); print(l["config5"]["per_leg"]); print(l["e2e"]["value"], l["clocks"])
PY
