#!/bin/bash
# Round 2, GPU call 20 (1 GPU): full GPU suite on the current build, kmeans after the early exit, bench.py with default flags
# (the driver's command) + the reference arm.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call20_pytest_all.txt 2>&1
tail -6 $O/r02_call20_pytest_all.txt
B=sycl-bench_b200/bin
rm -f $O/r02_kmeans_v3.csv
for s in 1048576 268435456; do timeout 600 $B/kmeans --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_kmeans_v3.csv > /dev/null 2>> $O/r02_call20.err; done
python - <<'PY'
hdr = None
for l in open('gpurun_out/r02_kmeans_v3.csv').read().splitlines():
    c = l.split(',')
    if l.startswith('# Benchmark name'):
        hdr = c
    elif hdr:
        d = dict(zip(hdr, c))
        print(c[0], d.get('problem-size'), 'kernel-time-min', d.get('kernel-time-min'), 'GiB', d.get('throughput-metric'), d.get('Verification'))
PY
( time timeout 900 python bench.py ) > $O/r02_bench_n1_final.json 2> $O/r02_bench_n1_final.err
tail -c 1500 $O/r02_bench_n1_final.json | head -c 1500; echo
python - <<'PY'
import json
l = json.loads(open('gpurun_out/r02_bench_n1_final.json').read().strip().splitlines()[0])
print({k: l[k] for k in ("value", "ms_per_step", "verified", "gpu_launches")}, l["roofline"]["frac"], l["roofline"]["per_leg"])
print(l["replicas"]["config3"]["per_leg"]); print(l["replicas"]["config4"]["per_leg"]); print(l["e2e"]["value"], l["clocks"])
PY
