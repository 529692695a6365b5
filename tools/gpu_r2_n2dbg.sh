#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
(while true; do free -g | sed -n 2p; sleep 3; done) > gpurun_out/mem_n$N.txt 2>&1 &
MP=$!
SGPU_DIST_DEBUG=1 SGPU_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 5 > gpurun_out/bench_c4_n$N.log 2> gpurun_out/bench_c4_n$N.err
echo "full rc=$?"; kill $MP
grep -E "dist\]|bench\]|Assertion|Signal" gpurun_out/bench_c4_n$N.err | head -30; cat gpurun_out/bench_c4_n$N.log | cut -c1-3000
tail -25 gpurun_out/mem_n$N.txt | awk '{print $3, $7}' | tr '\n' ';'
dmesg 2>/dev/null | grep -i -E "oom|killed" | tail -5
