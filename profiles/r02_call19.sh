#!/bin/bash
# Round 2, GPU call 19 (1 GPU): single-pass tiles for dense explicit flags (parity incl. both density verdicts, A/B by env over
# head densities), kmeans with 4 vectors per thread.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_single_kernel_widened.py tests/test_parity_gpu.py -q -m gpu -k "kmeans or segscan or flags or binaries" --maxfail=10 ) > $O/r02_call19_pytest.txt 2>&1
tail -4 $O/r02_call19_pytest.txt
: > $O/r02_ops_v15.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v15.jsonl 2>> $O/r02_call19.err; }
OPS=segscan_flags256,segscan_flagsrand32,segscan_flagsrand,segscan_flagsrand1024,segscan_flagsrand4096,segscan_flagsrand65536
run --ops $OPS
SB_SCAN_FLAGS_SP=0 run --ops $OPS
python profiles/fmt_jsonl.py $O/r02_ops_v15.jsonl
B=sycl-bench_b200/bin
rm -f $O/r02_kmeans_v2.csv
for s in 1048576 268435456; do timeout 600 $B/kmeans --size=$s --local=256 --num-runs=10 --warmup-run --no-verification --output=$O/r02_kmeans_v2.csv > /dev/null 2>> $O/r02_call19.err; done
python - <<'PY'
hdr = None
for l in open('gpurun_out/r02_kmeans_v2.csv').read().splitlines():
    c = l.split(',')
    if l.startswith('# Benchmark name'):
        hdr = c
    elif hdr:
        d = dict(zip(hdr, c))
        print(c[0], d.get('problem-size'), 'kernel-time-min', d.get('kernel-time-min'), 'GiB', d.get('throughput-metric'))
PY
