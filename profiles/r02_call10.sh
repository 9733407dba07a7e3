#!/bin/bash
# Round 2, GPU call 10 (1 GPU): which of the two changes slowed the plain look-back scan (A/B builds of sb_scan.cu).
mkdir -p gpurun_out
O=gpurun_out
: > $O/r02_ops_v6.jsonl
for lib in "" sycl-bench_b200/lib/tune/libsb_cuda_nowp.so sycl-bench_b200/lib/tune/libsb_cuda_noexcl.so sycl-bench_b200/lib/tune/libsb_cuda_neither.so sycl-bench_b200/lib/tune/libsb_cuda_prev.so; do
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan --dtypes i32,i64,f64 >> $O/r02_ops_v6.jsonl 2>> $O/r02_call10.err
done
python profiles/fmt_jsonl.py $O/r02_ops_v6.jsonl
