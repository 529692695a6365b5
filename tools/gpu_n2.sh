#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29533 tests/nccl_check.py > gpurun_out/nccl_check_n$N.log 2>&1; tail -6 gpurun_out/nccl_check_n$N.log
timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; tail -1 gpurun_out/bench_n$N.log
timeout 600 $TR --master-port 29535 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log
