#!/bin/bash
# Round 2, GPU call 1 (1 GPU): L2-residency probe + baseline op rates + ncu --set full of the kernels furthest from roofline.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r02_call1_smi.txt 2>&1
timeout 300 ./profiles/probes/l2_probe > $O/r02_l2_probe.txt 2>&1
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
for v in "0 0 0" "2 0 1" "2 1 1" "4 0 1" "4 1 1" "8 1 1"; do
  tag=$(echo $v | tr ' ' '_')
  timeout 300 ncu --metrics $M --clock-control none -s 2 -c 1 --csv --log-file $O/r02_l2_probe_ncu_$tag.csv ./profiles/probes/l2_probe $v > /dev/null 2>&1
done
timeout 900 python profiles/tune_scan.py --log2n 28 --iters 10 > $O/r02_ops_baseline.jsonl 2> $O/r02_ops_baseline.err
cap() { # name kernel-regex op dtype
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/r02_prof_$1 \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $3 --dtypes $4 > $O/r02_prof_$1.log 2>&1
  # the .ncu-rep files are ~20 MB each and gpurun_out is capped at 64 MiB: export the pages here, drop the report
  ncu -i $O/r02_prof_$1.ncu-rep --page raw --csv > $O/r02_prof_$1.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page source --csv > $O/r02_prof_$1.source.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page details > $O/r02_prof_$1.details.txt 2>/dev/null
  rm -f $O/r02_prof_$1.ncu-rep
}
cap flags_i32 scan_ws segscan_flags256 i32
cap flags_f64 scan_ws segscan_flags256 f64
cap lb65536_i32 scan_ws segscan_L65536 i32
cap lb65536_f64 scan_ws segscan_L65536 f64
cap segred64_i32 segreduce_rows segreduce64 i32
cap segred64_f64 segreduce_rows segreduce64 f64
du -sh $O; ls -la $O | tail -30
