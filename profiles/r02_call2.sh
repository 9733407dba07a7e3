#!/bin/bash
# Round 2, GPU call 2 (1 GPU): parity of the two-phase (block-cyclic) scan at world size 1, its rate against the look-back
# kernel for look-ahead depth / L2-policy variants, DRAM traffic under ncu.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cyclic_gpu.py -x -q -m gpu > $O/r02_call2_pytest.txt 2>&1
tail -5 $O/r02_call2_pytest.txt
: > $O/r02_cyclic_variants.jsonl
timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_cyclic --dtypes i32,i64,f32,f64 >> $O/r02_cyclic_variants.jsonl 2>> $O/r02_call2.err
for D in 1 2 3 4 6; do for H in 1 0; do
  SB_CYCLIC_DEPTH=$D SB_CYCLIC_HINT=$H timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants.jsonl 2>> $O/r02_call2.err
done; done
SB_CYCLIC_DEPTH=2 timeout 300 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops scan,scan_cyclic --dtypes i32,f32 >> $O/r02_cyclic_variants.jsonl 2>> $O/r02_call2.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for D in 2 3; do
  SB_CYCLIC_DEPTH=$D timeout 300 ncu --metrics $M --clock-control none -k regex:scan_cyclic -s 3 -c 1 --csv --log-file $O/r02_cyclic_ncu_D$D.csv \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops scan_cyclic --dtypes i64 > /dev/null 2>&1
done
du -sh $O
