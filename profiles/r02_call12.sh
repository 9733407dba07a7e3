#!/bin/bash
# Round 2, GPU call 12 (1 GPU): totals-based look-back in the plain scan kernel — full GPU suite, A/B against the chained round
# prefix, 2^30 sizes, DRAM traffic of the bench kernels under ncu.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call12_pytest_all.txt 2>&1
tail -6 $O/r02_call12_pytest_all.txt
: > $O/r02_ops_v8.jsonl
for lib in "" sycl-bench_b200/lib/tune/libsb_cuda_chain.so; do
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_cyclic >> $O/r02_ops_v8.jsonl 2>> $O/r02_call12.err
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops scan --dtypes i32,f32,i64 >> $O/r02_ops_v8.jsonl 2>> $O/r02_call12.err
done
SB_CYCLIC_DEPTH=1 timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic >> $O/r02_ops_v8.jsonl 2>> $O/r02_call12.err
python profiles/fmt_jsonl.py $O/r02_ops_v8.jsonl
M=dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for c in "scan scan_ws i64" "scan scan_ws f32" "reduce reduce_kernel i64" "reduce reduce_kernel f32" "scan_cyclic scan_cyclic i64"; do
  set -- $c
  timeout 300 ncu --metrics $M --clock-control none -k regex:$2 -s 3 -c 1 --csv --log-file $O/r02_traffic_$1_$3.csv \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $1 --dtypes $3 > /dev/null 2>&1
done
grep -h "dram__bytes\|time_duration" $O/r02_traffic_*.csv | awk -F'","' '{print $5, $(NF-2), $NF}' | cut -c1-200
