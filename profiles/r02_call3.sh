#!/bin/bash
# Round 2, GPU call 3 (1 GPU): two-phase scan with phase B loaded from L2 straight into registers; full GPU test suite; bench.py N=1.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cyclic_gpu.py -x -q -m gpu > $O/r02_call3_pytest_cyclic.txt 2>&1
tail -3 $O/r02_call3_pytest_cyclic.txt
: > $O/r02_cyclic_variants2.jsonl
for D in 1 2 3; do
  SB_CYCLIC_DEPTH=$D timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan,scan_cyclic --dtypes i32,i64,f32,f64 >> $O/r02_cyclic_variants2.jsonl 2>> $O/r02_call3.err
done
SB_CYCLIC_DEPTH=2 SB_CYCLIC_HINT=0 timeout 300 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops scan_cyclic --dtypes f32,i64 >> $O/r02_cyclic_variants2.jsonl 2>> $O/r02_call3.err
SB_CYCLIC_DEPTH=2 timeout 300 python profiles/tune_scan.py --worker --log2n 30 --iters 5 --ops scan,scan_cyclic --dtypes i32,f32 >> $O/r02_cyclic_variants2.jsonl 2>> $O/r02_call3.err
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum
for D in 2 3; do
  SB_CYCLIC_DEPTH=$D timeout 300 ncu --metrics $M --clock-control none -k regex:scan_cyclic -s 3 -c 1 --csv --log-file $O/r02_cyclic2_ncu_D$D.csv \
    python profiles/tune_scan.py --worker --log2n 28 --iters 1 --ops scan_cyclic --dtypes i64 > /dev/null 2>&1
done
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > $O/r02_call3_pytest_all.txt 2>&1
tail -5 $O/r02_call3_pytest_all.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r02_bench_n1_a.json 2> $O/r02_bench_n1_a.err
tail -c 600 $O/r02_bench_n1_a.json; tail -3 $O/r02_bench_n1_a.err
du -sh $O
