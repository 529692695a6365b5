#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/nccl_check_n$N.log 2>&1; tail -12 gpurun_out/nccl_check_n$N.log | cut -c1-400
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 5 --e2e-steps 0 --cpu-seconds 0.5 ${@:2} 2> gpurun_out/sb.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('${*:2}', round(j['ms_per_step'],3), j['phases_ms'], j['config']['parallelism'][40:120])" || tail -5 gpurun_out/sb.err; }
run 29551 --setbuild sharded
run 29552 --setbuild sharded --pull-lists
run 29553 --setbuild replicated
