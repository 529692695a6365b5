#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
run() { echo "== $*"; env "$@" timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1; }
dram() { env "$@" timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,smsp__inst_executed.sum --clock-control none --profile-from-start off -k regex:fastq_fused -s 1 -c 1 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep -E "dram__|lts__|smsp__inst" | tr -s ' ' | tr '\n' ';'; echo; }
run SGPU_VARIANT=
for v in p200 p1000 h128 lp2k lp16k pc1k nocopy noprobe; do run SGPU_VARIANT=$v; done
for f in 0 0.5 2 4; do run SGPU_VARIANT= SGPU_FUSED_PF=$f; done
dram SGPU_VARIANT=nocopy
dram SGPU_VARIANT=noprobe
