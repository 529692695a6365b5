#!/bin/bash
# round 2: the C4 strong-scaling bench on N GPUs of one box (N = $1), after a quick single-GPU sanity pass
N=${1:-2}
mkdir -p gpurun_out
(nvidia-smi topo -m; free -g; nproc; cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/memory.current 2>/dev/null; ulimit -l) > gpurun_out/topo_n$N.txt 2>&1
true
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --steps 5 > gpurun_out/bench_c4_n1.log 2> gpurun_out/bench_c4_n1.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 > gpurun_out/bench_c4_n$N.log 2> gpurun_out/bench_c4_n$N.err
fi
echo "bench rc=$?"; tail -8 gpurun_out/bench_c4_n$N.err; cat gpurun_out/bench_c4_n$N.log
