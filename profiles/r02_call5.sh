#!/bin/bash
# Round 2, GPU call 5 (1 GPU): tile / ring variants of the two-phase scan, A/B of the segmented-scan and segmented-reduction
# changes, full GPU test suite, bench.py N=1.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cyclic_gpu.py -x -q -m gpu > $O/r02_call5_pytest_cyclic.txt 2>&1
tail -3 $O/r02_call5_pytest_cyclic.txt
: > $O/r02_cyclic_variants3.jsonl
for lib in sycl-bench_b200/lib/tune/libsb_cuda_cy_*.so; do
  for D in 2 3; do
    SB_CUDA_LIB=$lib SB_CYCLIC_DEPTH=$D timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants3.jsonl 2>> $O/r02_call5.err
  done
done
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_cy_18_4_6_0.so SB_CYCLIC_DEPTH=4 timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants3.jsonl 2>> $O/r02_call5.err
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_cy_18_4_6_0.so SB_CYCLIC_DEPTH=6 timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants3.jsonl 2>> $O/r02_call5.err
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_cy_18_8_2_1.so SB_CYCLIC_DEPTH=1 timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants3.jsonl 2>> $O/r02_call5.err
# A/B: segmented look-back scans (lazy tiles, 14-worker 64-bit instantiation), segmented reduction (TMA ring)
OPS=segscan_L65536,segscan_flags256,segreduce64,segreduce256,segreduce1024,scan_group256
timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops $OPS > $O/r02_ops_v1.jsonl 2>> $O/r02_call5.err
SB_SCAN_LB64_NW=16 SB_SEGREDUCE_RING=0 timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops $OPS > $O/r02_ops_v1_off.jsonl 2>> $O/r02_call5.err
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=20 ) > $O/r02_call5_pytest_all.txt 2>&1
tail -15 $O/r02_call5_pytest_all.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1_c.json 2> $O/r02_bench_n1_c.err
tail -c 300 $O/r02_bench_n1_c.json; tail -3 $O/r02_bench_n1_c.err
du -sh $O
