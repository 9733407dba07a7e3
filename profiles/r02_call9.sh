#!/bin/bash
# Round 2, GPU call 9 (1 GPU): A/B of commit "per-warp prefixes + compile-time exclusive" against the commit before it, with
# unchanged kernels (reduce, iota, group scan) measured alongside as a control for the box.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > $O/r02_call9_smi.txt 2>&1
: > $O/r02_ops_v5.jsonl
for rep in 1 2; do
  timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops reduce,iota,scan_group256,scan,scan_cyclic --dtypes i32,i64 >> $O/r02_ops_v5.jsonl 2>> $O/r02_call9.err
  SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_prev.so timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops reduce,iota,scan_group256,scan,scan_cyclic --dtypes i32,i64 >> $O/r02_ops_v5.jsonl 2>> $O/r02_call9.err
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv >> $O/r02_call9_smi.txt 2>&1
python profiles/fmt_jsonl.py $O/r02_ops_v5.jsonl
cat $O/r02_call9_smi.txt
