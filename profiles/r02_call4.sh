#!/bin/bash
# Round 2, GPU call 4 (2 GPUs): block-cyclic scan across GPUs — single-process group (sb_comm, peer access) and one
# process per GPU (CUDA IPC mailboxes) — replicas, pattern binaries with --num-gpus=2, bench.py N=2 per strategy.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/r02_call4_topo.txt 2>&1
timeout 900 python -m pytest tests/test_cyclic_gpu.py tests/test_comm_gpu.py -x -q -m gpu > $O/r02_call4_pytest.txt 2>&1
tail -4 $O/r02_call4_pytest.txt
run() { # strategy extra...
  S=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 10 --warmup 3 \
     --scan-strategy $S "$@" 2> $O/r02_bench_n2_$S.err | tail -1 > $O/r02_bench_n2_$S.json
  tail -c 400 $O/r02_bench_n2_$S.json; echo; tail -2 $O/r02_bench_n2_$S.err
}
run cyclic
run reduce_then_scan --no-e2e --no-replicas
for D in 1 2 3 4; do
  SB_CYCLIC_DEPTH=$D timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 10 --warmup 3 \
     --no-e2e --no-replicas --no-config5 2>> $O/r02_call4.err | tail -1 > $O/r02_bench_n2_cyclic_D$D.json
done
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > $O/r02_bench_n1_b.json 2>> $O/r02_call4.err
du -sh $O
