#!/bin/bash
# round 2, run B: GPU tests, C4-like probe timing, the C4 bench on one GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
timeout 300 python tools/c4_probe.py 2>&1 | tail -2
SGPU_IDSET_BULK_MIN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or paf or txt or reads or sam or bam or diff" 2>&1 | tail -3
timeout 900 python bench.py --steps 5 > gpurun_out/bench_c4.log 2> gpurun_out/bench_c4.err
echo "bench rc=$?"; tail -5 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.log
