#!/bin/bash
# Round 2, GPU call 37 (1 GPU): ncu --set full of the two kernels of sb_linreg_error_f32 at 2^26 (which pipe bounds the fp32 row?).
mkdir -p gpurun_out
O=gpurun_out
B=sycl-bench_b200/bin
for k in fused2_kernel linreg_error_map; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o $O/r02_prof_lre_$k \
    $B/lin_reg_error --size=67108864 --local=256 --num-runs=2 --no-verification > $O/r02_prof_lre_$k.log 2>&1
  ncu -i $O/r02_prof_lre_$k.ncu-rep --page raw --csv > $O/r02_prof_lre_$k.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_lre_$k.ncu-rep --page source --csv > $O/r02_prof_lre_$k.source.csv 2>/dev/null
  rm -f $O/r02_prof_lre_$k.ncu-rep
done
ls -la $O | grep lre
