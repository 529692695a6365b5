#!/bin/bash
# quick iteration: GPU parity tests, device timing at 5M pairs (single + split output), ONT timing, phase timing, one ncu capture
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 > gpurun_out/quick_time.log 2>&1; tail -2 gpurun_out/quick_time.log
timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 --split > gpurun_out/quick_time_split.log 2>&1; tail -2 gpurun_out/quick_time_split.log
timeout 200 python tools/prof_step.py --ont 200000 --steps 2 > gpurun_out/quick_time_ont.log 2>&1; tail -2 gpurun_out/quick_time_ont.log
if [ -f scrubby_b200/lib/libscrubby_gpu_timing.so ]; then
SGPU_VARIANT=timing timeout 200 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep phases | tail -1
fi
for pf in $SGPU_PF_SWEEP; do echo "pf=$pf"; SGPU_FUSED_PF=$pf timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 2>&1 | grep "fused kernel" | tail -1; done
if [ "$1" != "noncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fastq_fused -s 1 -c 1 -f -o gpurun_out/fused python tools/prof_step.py --pairs 2000000 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
fi
