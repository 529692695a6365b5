#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1; }
for v in "$@"; do run SGPU_VARIANT=$v; done
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fastq_fused -s 1 -c 1 -f -o gpurun_out/fused python tools/prof_step.py --pairs 5000000 --steps 1 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
