#!/bin/bash
# one gpurun call: device timing of the fused kernel for the build variants given as arguments (tools/build_variant.sh),
# then the GPU parity suite against the variants named in $SGPU_SWEEP_TEST
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
run() { r=$(env SGPU_DEBUG=1 SGPU_VARIANT=$1 timeout 150 python tools/prof_step.py --pairs ${SGPU_SWEEP_PAIRS:-5000000} --steps 3 2>&1 | grep -E "fused kernel|Error|error|CTAs per SM" | tail -2 | tr '\n' ' '); echo "$1: $r" | tee -a gpurun_out/sweep.txt; }
run ""
for v in "$@"; do run $v; done
for v in $SGPU_SWEEP_TEST; do
  SGPU_VARIANT=$v timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$v.log 2>&1; echo "pytest $v: $(tail -1 gpurun_out/pytest_$v.log)" | tee -a gpurun_out/sweep.txt
done
for v in $SGPU_SWEEP_TIMING; do
  echo "$v: $(SGPU_VARIANT=$v timeout 150 python tools/prof_step.py --pairs 5000000 --steps 1 2>&1 | grep phases | tail -1)" | tee -a gpurun_out/sweep.txt
done
