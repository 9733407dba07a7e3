#!/bin/bash
# Round 2, GPU call 33 (1 GPU): 32-bit implicit heads on the look-back path: 14 workers x 10 rows (128 registers), with the
# per-row and the per-chunk phase A, against the default 18 x 8.
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_ops_v24.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v24.jsonl 2>> $O/r02_call33.err; }
for lib in "" lb32_14x10 lb32_14x10_chunk; do
  L=""; [ -n "$lib" ] && L=sycl-bench_b200/lib/tune/libsb_cuda_$lib.so
  SB_CUDA_LIB=$L run --ops segscan_L65536,segscan_L1M,segscan_L50000 --dtypes i32,f32
done
python profiles/fmt_jsonl.py $O/r02_ops_v24.jsonl
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_lb32_14x10_chunk.so timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "segscan_implicit or scan_group" --maxfail=5 2>&1 | tail -2
