#!/bin/bash
# Round 2, GPU call 8 (1 GPU): per-warp prefixes from the look-back warp + compile-time exclusive flag in the plain scans:
# full GPU suite, rates of the look-back and the two-phase kernel.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 2000 python -m pytest tests -q -m gpu --maxfail=10 ) > $O/r02_call8_pytest_all.txt 2>&1
tail -6 $O/r02_call8_pytest_all.txt
timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_cyclic > $O/r02_ops_v4.jsonl 2>> $O/r02_call8.err
timeout 600 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops scan,scan_cyclic --dtypes i32,f32 >> $O/r02_ops_v4.jsonl 2>> $O/r02_call8.err
cat $O/r02_ops_v4.jsonl | python profiles/fmt_jsonl.py /dev/stdin 2>/dev/null || cat $O/r02_ops_v4.jsonl
