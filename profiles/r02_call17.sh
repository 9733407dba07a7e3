#!/bin/bash
# Round 2, GPU call 17 (1 GPU): full GPU suite with the widened single-kernel rows (lin_reg_error, kmeans), their drop-in
# binaries at upstream's suite sizes and at HBM-sized inputs, segmented scans after the dispatch changes.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call17_pytest_all.txt 2>&1
tail -6 $O/r02_call17_pytest_all.txt
B=sycl-bench_b200/bin
rm -f $O/r02_single_kernel_widened.csv
for s in 65536 67108864; do timeout 600 $B/lin_reg_error --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_single_kernel_widened.csv > /dev/null 2>> $O/r02_call17.err; done
for s in 1048576 268435456; do timeout 600 $B/kmeans --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_single_kernel_widened.csv > /dev/null 2>> $O/r02_call17.err; done
python - <<'PY'
import sys
sys.path.insert(0, '.')
from oracle import oracle as O
rows = {}
import csv
for line in open('gpurun_out/r02_single_kernel_widened.csv'):
    pass
txt = open('gpurun_out/r02_single_kernel_widened.csv').read().splitlines()
hdr = None
for l in txt:
    c = l.split(',')
    if l.startswith('# Benchmark name'):
        hdr = c
    elif hdr:
        d = dict(zip(hdr, c))
        print(c[0], d.get('problem-size'), 'kernel-time-min', d.get('kernel-time-min'), 'throughput', d.get('throughput-metric'), d.get('Verification'))
PY
: > $O/r02_ops_v13.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v13.jsonl 2>> $O/r02_call17.err; }
run --ops segscan_L16384,segscan_flags256,segscan_flagsrand,scan_group256
run --ops segscan_flags256,scan_group256 --log2n 30 --dtypes i32,f32,i64
python profiles/fmt_jsonl.py $O/r02_ops_v13.jsonl
