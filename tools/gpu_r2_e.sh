#!/bin/bash
# round 2, run E: evidence runs on one GPU -- long ids, ncu captures of the non-fused kernels, compute-sanitizer
mkdir -p gpurun_out
timeout 300 python tools/long_ids.py > gpurun_out/r02_long_ids.txt 2>&1; tail -3 gpurun_out/r02_long_ids.txt | cut -c1-600
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"lines_tile|idset_count|idset_scatter|idset_page|idset_insert|fastq_copy|fastq_record" -c 8 -f -o gpurun_out/r02_other_kernels python tools/prof_other.py > gpurun_out/ncu_other.log 2>&1; tail -2 gpurun_out/ncu_other.log
bash tools/gpu_sanitize.sh
