#!/bin/bash
# tools/build_variant.sh <name> "<nvcc -D flags>": builds libscrubby_gpu_<name>.so and prints the fused kernel's registers / spills
name=$1; shift
SGPU_VARIANT=$name SGPU_NVCC_FLAGS="$*" python -m scrubby_b200.build --force -v 2>&1 | grep -A2 "fastq_fused_kernelILb0" | grep -E "registers|spill" | tr '\n' ' '
echo " <- $name: $*"
