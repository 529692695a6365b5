#!/bin/bash
# quick iteration: fused-kernel parity subset, device timing at 5M pairs, one ncu capture
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "fused or shard or golden_fastq or adversarial or c3 or c1" > gpurun_out/pytest_quick.log 2>&1; tail -2 gpurun_out/pytest_quick.log
timeout 300 python tools/prof_step.py --pairs 5000000 --steps 4 > gpurun_out/quick_time.log 2>&1; tail -3 gpurun_out/quick_time.log
timeout 300 python tools/prof_step.py --pairs 5000000 --steps 3 --split > gpurun_out/quick_time_split.log 2>&1; tail -1 gpurun_out/quick_time_split.log
if [ "$1" != "noncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fastq_fused -s 1 -c 1 -f -o gpurun_out/fused python tools/prof_step.py --pairs 2000000 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
fi
