#!/bin/bash
# Round 2, GPU call 24 (1 GPU): 8-row tiles for 64-bit explicit flags (A/B against lib/tune/libsb_cuda_flags_k6.so) + parity;
# ncu --set full of the kernels changed in the second half of the round (evidence for DESIGN.md 4).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "segscan or flags or config3" --maxfail=10 ) > $O/r02_call24_pytest.txt 2>&1
tail -4 $O/r02_call24_pytest.txt
: > $O/r02_ops_v18.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v18.jsonl 2>> $O/r02_call24.err; }
OPS=segscan_flags256,segscan_flagsrand32,segscan_flagsrand,segscan_flagsrand4096
run --ops $OPS --dtypes i64,f64
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_flags_k6.so run --ops $OPS --dtypes i64,f64
run --ops segscan_flags256 --dtypes i64,f64 --log2n 30
python profiles/fmt_jsonl.py $O/r02_ops_v18.jsonl
cap() { # name kernel-regex op dtype
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/r02_prof_$1 \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $3 --dtypes $4 > $O/r02_prof_$1.log 2>&1
  ncu -i $O/r02_prof_$1.ncu-rep --page raw --csv > $O/r02_prof_$1.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page source --csv > $O/r02_prof_$1.source.csv 2>/dev/null
  rm -f $O/r02_prof_$1.ncu-rep
}
cap final_flags256_i64 scan_independent segscan_flags256 i64
cap final_flagsrand_i32 scan_ws segscan_flagsrand i32
cap final_flagsrand_i64 scan_ws segscan_flagsrand i64
cap final_L65536_i32 scan_ws segscan_L65536 i32
cap final_scan_i64 scan_ws scan i64
B=sycl-bench_b200/bin
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kmeans_kernel -s 3 -c 1 -f -o $O/r02_prof_final_kmeans_f64 \
  $B/kmeans --size=134217728 --local=256 --num-runs=2 --no-verification > $O/r02_prof_final_kmeans_f64.log 2>&1
ncu -i $O/r02_prof_final_kmeans_f64.ncu-rep --page raw --csv > $O/r02_prof_final_kmeans_f64.raw.csv 2>/dev/null
ncu -i $O/r02_prof_final_kmeans_f64.ncu-rep --page source --csv > $O/r02_prof_final_kmeans_f64.source.csv 2>/dev/null
rm -f $O/r02_prof_final_kmeans_f64.ncu-rep
du -sh $O
