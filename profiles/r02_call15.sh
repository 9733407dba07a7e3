#!/bin/bash
# Round 2, GPU call 15 (1 GPU): (a) explicit flags: check kernel + gated kernel pair (independent tiles when every tile
# starts on a head); (b) segmented reduction: lane-owns-128-bytes arithmetic in the ring kernel.  Parity first, then A/B
# (env switches on the same build), then ncu --set full of the segmented-reduce ring kernel, int32 and int16 at L=64.
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "scan or seg" --maxfail=10 ) > $O/r02_call15_pytest.txt 2>&1
tail -4 $O/r02_call15_pytest.txt
: > $O/r02_ops_v11.jsonl
run() { timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 "$@" >> $O/r02_ops_v11.jsonl 2>> $O/r02_call15.err; }
run --ops segscan_flags256,segscan_flagsrand
SB_SCAN_FLAGS_GATE=0 run --ops segscan_flags256,segscan_flagsrand
run --ops segreduce64,segreduce128,segreduce256 --dtypes i16,i32,f32,i64,f64
SB_SEGREDUCE_LANE=0 run --ops segreduce64,segreduce128,segreduce256 --dtypes i16,i32,f32,i64,f64
python profiles/fmt_jsonl.py $O/r02_ops_v11.jsonl
cap() { # name kernel-regex op dtype
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o $O/r02_prof_$1 \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops $3 --dtypes $4 > $O/r02_prof_$1.log 2>&1
  ncu -i $O/r02_prof_$1.ncu-rep --page raw --csv > $O/r02_prof_$1.raw.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page source --csv > $O/r02_prof_$1.source.csv 2>/dev/null
  ncu -i $O/r02_prof_$1.ncu-rep --page details > $O/r02_prof_$1.details.txt 2>/dev/null
  rm -f $O/r02_prof_$1.ncu-rep
}
cap segring64_i32_lane segreduce_ring segreduce64 i32
SB_SEGREDUCE_LANE=0 cap segring64_i32_rows segreduce_ring segreduce64 i32
cap segring64_i16_lane segreduce_ring segreduce64 i16
cap flags256_i32_gate scan_independent segscan_flags256 i32
du -sh $O
