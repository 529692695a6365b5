#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_quick.log 2>&1; tail -3 gpurun_out/pytest_quick.log
run() { echo "== $*"; env "$@" timeout 200 python tools/prof_step.py --pairs 5000000 --steps 3 2>&1 | grep -E "fused kernel|Error|error" | tail -1; }
run SGPU_VARIANT=
for v in "$@"; do run SGPU_VARIANT=$v; done
echo split; timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 --split 2>&1 | grep -E "fused kernel|Error|error" | tail -1
echo ont; timeout 200 python tools/prof_step.py --ont 200000 --steps 2 2>&1 | grep -E "fused kernel|Error|error" | tail -1
SGPU_VARIANT=timing SGPU_FUSED_TRACE=gpurun_out/trace.bin timeout 200 python tools/prof_step.py --pairs 5000000 --steps 2 2>&1 | grep -E "phases" | tail -1
