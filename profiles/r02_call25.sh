#!/bin/bash
# Round 2, GPU call 25 (1 GPU): 32-bit explicit flags on 14 workers x 8 rows (lib/tune/libsb_cuda_flags32_14x8.so) against 16 x 6.
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_ops_v19.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v19.jsonl 2>> $O/r02_call25.err; }
OPS=segscan_flags256,segscan_flagsrand32,segscan_flagsrand,segscan_flagsrand4096
run --ops $OPS --dtypes i32,f32
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_flags32_14x8.so run --ops $OPS --dtypes i32,f32
run --ops segscan_flags256 --dtypes i32,f32 --log2n 30
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_flags32_14x8.so run --ops segscan_flags256 --dtypes i32,f32 --log2n 30
python profiles/fmt_jsonl.py $O/r02_ops_v19.jsonl
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_flags32_14x8.so timeout 600 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "flags" --maxfail=5 2>&1 | tail -2
