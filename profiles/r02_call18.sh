#!/bin/bash
# Round 2, GPU call 18 (1 GPU): widened rows after the kmeans staging fix (tests, binaries, per-kernel times under ncu),
# 18-worker tiles for the 32-bit implicit look-back (A/B by env), run-suite quicktest.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_single_kernel_widened.py tests/test_harness.py tests/test_parity_gpu.py -q -m gpu -k "widened or quicktest or segscan or lin_reg or kmeans or binaries" --maxfail=10 ) > $O/r02_call18_pytest.txt 2>&1
tail -4 $O/r02_call18_pytest.txt
B=sycl-bench_b200/bin
rm -f $O/r02_single_kernel_widened.csv
for s in 65536 67108864; do timeout 600 $B/lin_reg_error --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_single_kernel_widened.csv > /dev/null 2>> $O/r02_call18.err; done
for s in 1048576 268435456; do timeout 600 $B/kmeans --size=$s --local=256 --num-runs=10 --warmup-run --output=$O/r02_single_kernel_widened.csv > /dev/null 2>> $O/r02_call18.err; done
python - <<'PY'
hdr = None
for l in open('gpurun_out/r02_single_kernel_widened.csv').read().splitlines():
    c = l.split(',')
    if l.startswith('# Benchmark name'):
        hdr = c
    elif hdr:
        d = dict(zip(hdr, c))
        print(c[0], d.get('problem-size'), 'kernel-time-min', d.get('kernel-time-min'), 'GiB', d.get('throughput-metric'), d.get('Verification'))
PY
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_launches_lin_reg_error.csv \
  $B/lin_reg_error --size=67108864 --local=256 --num-runs=2 --no-verification > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $O/r02_launches_kmeans.csv \
  $B/kmeans --size=268435456 --local=256 --num-runs=2 --no-verification > /dev/null 2>&1
grep -E "kernel" $O/r02_launches_lin_reg_error.csv | cut -d, -f5,13- | cut -c1-200 | tail -18
grep -E "kernel" $O/r02_launches_kmeans.csv | cut -d, -f5,13- | cut -c1-200 | tail -12
: > $O/r02_ops_v14.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v14.jsonl 2>> $O/r02_call18.err; }
run --ops segscan_L65536,segscan_L1M,segscan_L50000 --dtypes i32,f32
SB_SCAN_LB32_NW=16 run --ops segscan_L65536,segscan_L1M,segscan_L50000 --dtypes i32,f32
python profiles/fmt_jsonl.py $O/r02_ops_v14.jsonl
