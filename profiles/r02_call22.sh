#!/bin/bash
# Round 2, GPU call 22 (1 GPU): phase A of the segmented scans with one warp sum per chunk instead of one per row (A/B against
# lib/tune/libsb_cuda_phasea_per_row.so) + parity of the segmented scans; ncu --set full of kmeans<double>.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "segscan or flags or config3" --maxfail=10 ) > $O/r02_call22_pytest.txt 2>&1
tail -4 $O/r02_call22_pytest.txt
: > $O/r02_ops_v16.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v16.jsonl 2>> $O/r02_call22.err; }
OPS=segscan_flagsrand32,segscan_flagsrand,segscan_flagsrand4096,segscan_L65536,segscan_L50000
run --ops $OPS
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_phasea_per_row.so run --ops $OPS
SB_SCAN_FLAGS_GATE=0 run --ops segscan_flags256
SB_SCAN_FLAGS_GATE=0 SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_phasea_per_row.so run --ops segscan_flags256
python profiles/fmt_jsonl.py $O/r02_ops_v16.jsonl
B=sycl-bench_b200/bin
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_kernel -s 3 -c 1 -f -o $O/r02_prof_kmeans_f64 \
  $B/kmeans --size=134217728 --local=256 --num-runs=2 --no-verification > $O/r02_prof_kmeans_f64.log 2>&1
ncu -i $O/r02_prof_kmeans_f64.ncu-rep --page raw --csv > $O/r02_prof_kmeans_f64.raw.csv 2>/dev/null
ncu -i $O/r02_prof_kmeans_f64.ncu-rep --page source --csv > $O/r02_prof_kmeans_f64.source.csv 2>/dev/null
ncu -i $O/r02_prof_kmeans_f64.ncu-rep --page details > $O/r02_prof_kmeans_f64.details.txt 2>/dev/null
rm -f $O/r02_prof_kmeans_f64.ncu-rep
