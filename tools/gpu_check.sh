#!/bin/bash
# one gpurun call: GPU parity tests, smoke, bench, ncu launch list, ncu full capture of the fused kernel
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -2 gpurun_out/bench.log
timeout 600 python bench.py --steps 5 --warmup 3 --split > gpurun_out/bench_split.log 2>&1; tail -1 gpurun_out/bench_split.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fastq_fused -s 1 -c 1 -f -o gpurun_out/fused python tools/prof_step.py --pairs 10000000 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log

timeout 300 python tools/config_times.py > gpurun_out/config_times.log 2>&1; grep -E "^C[1-5]" gpurun_out/config_times.log
timeout 300 python tools/general_time.py > gpurun_out/general_time.log 2>&1; tail -3 gpurun_out/general_time.log
timeout 300 python tools/bam_time.py > gpurun_out/bam_time.log 2>&1; tail -2 gpurun_out/bam_time.log
timeout 400 python tools/c4_full.py > gpurun_out/c4_full.log 2>&1; tail -1 gpurun_out/c4_full.log | cut -c1-400
