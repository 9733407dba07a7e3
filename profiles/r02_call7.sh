#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): fused all-reduce (sb_reduce_all) — tests, bench N=2 fused vs NCCL exchange; unchained defaults.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py tests/test_bench_contract.py tests/test_micro_dram.py -x -q -m gpu > $O/r02_call7_pytest.txt 2>&1
tail -4 $O/r02_call7_pytest.txt
for X in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --steps 20 --warmup 5 \
     --reduce-exchange $X --no-e2e --no-replicas 2> $O/r02_bench_n2_red_$X.err | tail -1 > $O/r02_bench_n2_red_$X.json
  tail -2 $O/r02_bench_n2_red_$X.err | head -1
done
SB_CYCLIC_DEPTH=2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29623 bench.py --gpus 2 --steps 20 --warmup 5 \
     --no-e2e --no-replicas 2>> $O/r02_call7.err | tail -1 > $O/r02_bench_n2_red_fused_D2.json
timeout 600 python profiles/tune_scan.py --worker --log2n 28 --iters 10 --ops segscan_L65536,segscan_flags256,scan_group256,segreduce64 > $O/r02_ops_v3.jsonl 2>> $O/r02_call7.err
du -sh $O
