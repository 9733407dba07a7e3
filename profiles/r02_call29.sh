#!/bin/bash
# Round 2, GPU call 29 (1 GPU): tile shapes of the two-phase (block-cyclic) scan kernel: 14 x 10 and 16 x 9 rows against 18 x 8.
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_ops_v23.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --iters 10 "$@" >> $O/r02_ops_v23.jsonl 2>> $O/r02_call29.err; }
for lib in "" cy_14_10_3_0 cy_16_9_3_0; do
  L=""; [ -n "$lib" ] && L=sycl-bench_b200/lib/tune/libsb_cuda_$lib.so
  SB_CUDA_LIB=$L run --ops scan_cyclic
done
python profiles/fmt_jsonl.py $O/r02_ops_v23.jsonl
