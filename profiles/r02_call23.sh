#!/bin/bash
# Round 2, GPU call 23 (1 GPU): kmeans with unguarded loads for full blocks; phase-A form selected per mode / type; full GPU suite.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call23_pytest_all.txt 2>&1
tail -4 $O/r02_call23_pytest_all.txt
B=sycl-bench_b200/bin
rm -f $O/r02_kmeans_v5.csv
for s in 1048576 268435456; do timeout 600 $B/kmeans --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_kmeans_v5.csv > /dev/null 2>> $O/r02_call23.err; done
grep -v "^#" $O/r02_kmeans_v5.csv | cut -d, -f1,2,8,15
: > $O/r02_ops_v17.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v17.jsonl 2>> $O/r02_call23.err; }
run --ops segscan_flagsrand,segscan_L65536,segscan_flags256,scan,scan_group256,segreduce64
python profiles/fmt_jsonl.py $O/r02_ops_v17.jsonl
