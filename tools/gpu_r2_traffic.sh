#!/bin/bash
# ncu --set full capture of ONE fused-kernel launch of the bench's own timed region (C4 on one GPU) -> DRAM traffic per launch
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:fastq_fused -c 1 -f -o gpurun_out/r02_fused_c4 python bench.py --steps 1 --e2e-steps 0 --cpu-seconds 0.2 > gpurun_out/ncu_fused_c4.log 2>&1
tail -3 gpurun_out/ncu_fused_c4.log | cut -c1-300; ls -la gpurun_out/r02_fused_c4.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file gpurun_out/r02_launches_c4_bench.csv python bench.py --steps 2 --e2e-steps 0 --cpu-seconds 0.2 > gpurun_out/launches_c4.log 2>&1; tail -1 gpurun_out/launches_c4.log | cut -c1-200
