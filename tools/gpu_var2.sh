#!/bin/bash
# device timing + DRAM bytes of build variants: tools/gpu_var2.sh v1 v2 ...   ("-" = the default build)
mkdir -p gpurun_out
for v in "$@"; do
  [ "$v" = "-" ] && v=""
  echo "variant=[$v]"
  SGPU_VARIANT=$v timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1
  SGPU_VARIANT=$v timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|smsp__inst" | tr -s ' ' | tr '\n' ';'; echo
done
