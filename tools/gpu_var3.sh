#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|L2 fetch|Error|error" | tail -2; }
dram() { env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|smsp__inst" | tr -s ' ' | tr '\n' ';'; echo; }
run SGPU_VARIANT=
run SGPU_VARIANT=f9
dram SGPU_VARIANT=f9
for g in 32 64 128; do run SGPU_VARIANT= SGPU_L2_FETCH=$g; dram SGPU_VARIANT= SGPU_L2_FETCH=$g; done
run SGPU_VARIANT=f9 SGPU_L2_FETCH=32
SGPU_VARIANT=timing timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 2>&1 | grep -E "phases" | tail -1
