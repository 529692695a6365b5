#!/bin/bash
# parity tests with the default build, then device timing of build variants: tools/gpu_var.sh v1 v2 ...
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
for v in "" "$@"; do
  echo "variant=[$v]"
  SGPU_VARIANT=$v timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1
  SGPU_VARIANT=$v timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 --split 2>&1 | grep -E "fused kernel|Error|error" | tail -1
  SGPU_VARIANT=$v timeout 200 python tools/prof_step.py --ont 200000 --steps 2 2>&1 | grep -E "fused kernel|Error|error" | tail -1
done
if [ -f scrubby_b200/lib/libscrubby_gpu_timing.so ]; then
SGPU_VARIANT=timing SGPU_FUSED_TRACE=gpurun_out/trace.bin timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 2>&1 | grep -E "phases" | tail -1
fi
SGPU_VARIANT=$NCU_VARIANT timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|gpu__time|smsp__inst"
