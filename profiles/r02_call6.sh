#!/bin/bash
# Round 2, GPU call 6 (1 GPU): where does the two-phase scan spend its time (ncu --set full, 64-bit vs 32-bit); A/B of the
# unchained row scans for the segmented look-back kernels.
mkdir -p gpurun_out
O=gpurun_out
cap() { # name kernel-regex op dtype
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/r02_prof_$1 \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $3 --dtypes $4 > $O/r02_prof_$1.log 2>&1
  ncu -i $O/r02_prof_$1.ncu-rep --page raw --csv > $O/r02_prof_$1.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page source --csv > $O/r02_prof_$1.source.csv 2>/dev/null
  rm -f $O/r02_prof_$1.ncu-rep
}
cap cyclic_i64 scan_cyclic scan_cyclic i64
cap cyclic_f32 scan_cyclic scan_cyclic f32
cap scanws_i64 scan_ws scan i64
OPS=segscan_L65536,segscan_flags256,scan_group256
timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops $OPS > $O/r02_ops_v2.jsonl 2>> $O/r02_call6.err
SB_CUDA_LIB=sycl-bench_b200/lib/tune/libsb_cuda_unchained.so timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops $OPS > $O/r02_ops_v2_unchained.jsonl 2>> $O/r02_call6.err
timeout 600 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops segscan_flags256,scan_group256,scan --dtypes i32,f32 > $O/r02_ops_v2_2p30.jsonl 2>> $O/r02_call6.err
du -sh $O
