#!/bin/bash
# round 2: the C4 strong-scaling bench on N GPUs of one box (N = $1), then the multi-GPU parity check
N=${1:-2}
mkdir -p gpurun_out
(nvidia-smi topo -m; free -g; nproc) > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 5 > gpurun_out/bench_c4_n1.log 2> gpurun_out/bench_c4_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 > gpurun_out/bench_c4_n$N.log 2> gpurun_out/bench_c4_n$N.err
fi
echo "bench rc=$?"; grep -E "bench\]|Error|Signal" gpurun_out/bench_c4_n$N.err | tail -8; cat gpurun_out/bench_c4_n$N.log
if [ "$N" != "1" ] && [ -z "$SKIP_CHECK" ]; then
  timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/nccl_check_n$N.log 2>&1; tail -3 gpurun_out/nccl_check_n$N.log
fi
