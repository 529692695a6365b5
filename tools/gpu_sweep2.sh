#!/bin/bash
# one gpurun call: tools/c4_probe.py for the default build and every variant given as argument
mkdir -p gpurun_out
: > gpurun_out/sweep2.txt
run() { r=$(env SGPU_VARIANT=$1 timeout 200 python tools/c4_probe.py ${SGPU_SWEEP_ARGS} 2>&1 | grep -E "variant|Error|error" | tail -1); echo "$r" | tee -a gpurun_out/sweep2.txt; }
run ""
for v in "$@"; do run $v; done
