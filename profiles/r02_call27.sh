#!/bin/bash
# Round 2, GPU call 27 (1 GPU): tile shapes of the global-scan instantiation with the same ~70 KB tile but fewer, fatter warps
# (14 x 10, 12 x 12 rows: 128 registers per thread) and 16 x 9 against the default 18 workers x 8 rows.
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_ops_v21.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v21.jsonl 2>> $O/r02_call27.err; }
for lib in "" wide14x10 wide12x12 wide16x9; do
  L=""; [ -n "$lib" ] && L=sycl-bench_b200/lib/tune/libsb_cuda_$lib.so
  SB_CUDA_LIB=$L run --ops scan
  SB_CUDA_LIB=$L run --ops scan --dtypes i32,i64 --log2n 30
done
python profiles/fmt_jsonl.py $O/r02_ops_v21.jsonl
