#!/bin/bash
# Round 2, GPU call 14 (1 GPU): head flags as bit masks (AND + add instead of test / select per element) and packed flag-byte
# decode in the segmented scans: scan parity subset + A/B against the previous build (lib/tune/libsb_cuda_prev.so).
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "scan or seg" --maxfail=10 ) > $O/r02_call14_pytest.txt 2>&1
tail -4 $O/r02_call14_pytest.txt
: > $O/r02_ops_v10.jsonl
for lib in "" sycl-bench_b200/lib/tune/libsb_cuda_prev.so; do
  SB_CUDA_LIB=$lib timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_group256,segscan_L65536,segscan_flags256 >> $O/r02_ops_v10.jsonl 2>> $O/r02_call14.err
done
python profiles/fmt_jsonl.py $O/r02_ops_v10.jsonl
