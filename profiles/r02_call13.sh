#!/bin/bash
# Round 2, GPU call 13 (1 GPU): publish-ahead in the chained look-back kernel (all modes): full GPU suite + A/B.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call13_pytest_all.txt 2>&1
tail -6 $O/r02_call13_pytest_all.txt
: > $O/r02_ops_v9.jsonl
for lib in "" sycl-bench_b200/lib/tune/libsb_cuda_nopa.so; do
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,segscan_L65536,segscan_flags256 >> $O/r02_ops_v9.jsonl 2>> $O/r02_call13.err
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops scan --dtypes i32,f32,i64 >> $O/r02_ops_v9.jsonl 2>> $O/r02_call13.err
done
python profiles/fmt_jsonl.py $O/r02_ops_v9.jsonl
