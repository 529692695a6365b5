#!/bin/bash
mkdir -p gpurun_out
N=2
run() { env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $N --steps 5 --e2e-steps 0 --cpu-seconds 0.5 2>/dev/null | python -c "import sys,json; j=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', j['ms_per_step'], j['phases_ms'])"; }
run A=1 29551
run SGPU_BENCH_NOSAMPLER=1 29552
run SGPU_BENCH_SYNC=1 29553
run SGPU_IDSET_BULK_MIN=1000000000000 29554
