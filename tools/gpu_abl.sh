#!/bin/bash
# ablation timings of the fused kernel (timing only, outputs of the ablated builds are garbage) + per-tile trace
mkdir -p gpurun_out
for v in "" noprobe nolb nocopy; do
  echo "variant=[$v]"
  SGPU_VARIANT=$v timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel" | tail -1
done
echo "timing variant"
SGPU_VARIANT=timing SGPU_FUSED_TRACE=gpurun_out/trace.bin timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 2>&1 | grep -E "phases|fused kernel" | tail -2
for v in "" noprobe; do
  SGPU_VARIANT=$v timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|gpu__time"
done
