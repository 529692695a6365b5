#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 5 --e2e-steps 0 --cpu-seconds 0.5 --gather $1 2> gpurun_out/gather_$1.err | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', j['ms_per_step'], j['phases_ms'], j['config']['parallelism'][:90])"; grep -E "bench\]" gpurun_out/gather_$1.err | head -3; }
run pull 29551
run nccl 29552
