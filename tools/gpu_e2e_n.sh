#!/bin/bash
# end-to-end arm at N GPUs with and without NUMA binding: tools/gpu_e2e_n.sh N
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" >> gpurun_out/topo.txt
for f in "" "--no-bind"; do
timeout 600 $TR --master-port 29544 bench.py --gpus $N --steps 5 --warmup 3 --cpu-seconds 1 $f 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$f', j['n_gpus'], round(j['value']/1e9,2), 'G reads/s', 'e2e', round(j['e2e']['value']/1e6,1), j['e2e']['ms_each_rank0'], j['config']['host_cpus_rank0'])"
done
