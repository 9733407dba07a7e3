#!/bin/bash
# Round 2, GPU call 32 (1 GPU): what the driver runs at round end on the final build — smoke(), the reference arm, bench.py.
mkdir -p gpurun_out
O=gpurun_out
( time python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) 2>&1 | tail -5
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
tail -c 900 $O/r02_bench_reference_arm.json; echo; tail -3 $O/r02_bench_reference_arm.err
