#!/bin/bash
# Round 2, GPU call 34 (1 GPU): the driver's round-end sequence on the final commit: full GPU suite, smoke(), bench.py.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_final_pytest_gpu.txt 2>&1
tail -4 $O/r02_final_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time timeout 900 python bench.py ) > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[0])
print({k: l[k] for k in ("value", "ms_per_step", "verified", "gpu_launches")}, l["roofline"]["frac"], l["roofline"]["per_leg"])
print(l["replicas"]["config3"]["per_leg"]); print(l["replicas"]["config4"]["per_leg"]); print(l["e2e"]["value"], l["clocks"])
PY
