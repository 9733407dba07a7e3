#!/bin/bash
# Round 2, GPU call 26 (1 GPU): 64-bit implicit look-back on 14 workers x 10 rows (70 KB tiles) against 14 x 8 (lib/tune/libsb_cuda_lb64_k8.so).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "segscan or scan_group or config3" --maxfail=5 2>&1 | tail -2
: > $O/r02_ops_v20.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v20.jsonl 2>> $O/r02_call26.err; }
OPS=segscan_L16384,segscan_L65536,segscan_L1M,segscan_L50000
run --ops $OPS --dtypes i64,f64
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_lb64_k8.so run --ops $OPS --dtypes i64,f64
python profiles/fmt_jsonl.py $O/r02_ops_v20.jsonl
