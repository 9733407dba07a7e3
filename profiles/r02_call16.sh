#!/bin/bash
# Round 2, GPU call 16 (1 GPU): segmented look-back that waits only up to the nearest head (SB_SCAN_SEG_EARLY) against the
# previous behaviour (lib/tune/libsb_cuda_noearly.so); multi-block tile-head check for the explicit-flag kernel pair.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "scan or seg" --maxfail=10 ) > $O/r02_call16_pytest.txt 2>&1
tail -4 $O/r02_call16_pytest.txt
: > $O/r02_ops_v12.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 "$@" >> $O/r02_ops_v12.jsonl 2>> $O/r02_call16.err; }
OPS=segscan_L16384,segscan_L65536,segscan_L1M,segscan_L50000,segscan_flags256,segscan_flagsrand
run --ops $OPS
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_noearly.so run --ops $OPS
SB_SCAN_FLAGS_GATE=0 run --ops segscan_flagsrand
python profiles/fmt_jsonl.py $O/r02_ops_v12.jsonl
cap() { # name kernel-regex op dtype
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/r02_prof_$1 \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $3 --dtypes $4 > $O/r02_prof_$1.log 2>&1
  ncu -i $O/r02_prof_$1.ncu-rep --page raw --csv > $O/r02_prof_$1.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page source --csv > $O/r02_prof_$1.source.csv 2>/dev/null
  rm -f $O/r02_prof_$1.ncu-rep
}
cap early_L65536_i32 scan_ws segscan_L65536 i32
cap early_flagsrand_i32 scan_ws segscan_flagsrand i32
