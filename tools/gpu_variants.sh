#!/bin/bash
# device timing (and DRAM bytes for b2 / b1) of build variants made with SGPU_VARIANT=<name> SGPU_NVCC_FLAGS="-D..." python -m scrubby_b200.build
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1; }
dram() { env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|smsp__inst" | tr -s ' ' | tr '\n' ';'; echo; }
run SGPU_VARIANT=
for v in "$@"; do run SGPU_VARIANT=$v; done
dram SGPU_VARIANT=b2
dram SGPU_VARIANT=b1
