#!/bin/bash
# round 2, run A: GPU tests + the C4 bench on one GPU
mkdir -p gpurun_out
(free -g; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv) > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 > gpurun_out/bench_c4.log 2> gpurun_out/bench_c4.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.log
