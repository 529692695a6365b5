"""N-rank diagnostic: the all-gather of the evidence shards and the set build from its output, separated by
synchronisation (torchrun --nproc-per-node N tools/n2_setbuild.py [pairs])."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import gen_txt_full
from scrubby_b200 import api
from scrubby_b200 import dist as sdist

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = api.Context(local)
full = gen_txt_full(pairs, dev)
total = int(full.numel())
per = sdist.evidence_shard_len(total, world)
d_ev = torch.zeros(per + 16, dtype=torch.uint8, device=dev)
mine = full[rank * per: min(total, (rank + 1) * per)]
d_ev[: mine.numel()] = mine
out = torch.empty(world * per + 16, dtype=torch.uint8, device=dev)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


for it in range(4):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = ev()
    g = sdist.replicate_file_dev(d_ev, per, total, dist if world > 1 else None, out)
    e1 = ev()
    torch.cuda.synchronize()
    t_ag = e0.elapsed_time(e1)
    same = bool(torch.equal(g, full))
    e2 = ev()
    s1 = api.IdSet.from_txt(ctx, g)
    e3 = ev()
    torch.cuda.synchronize()
    e4 = ev()
    s2 = api.IdSet.from_txt(ctx, full)
    e5 = ev()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s3 = api.IdSet.from_txt(ctx, full)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    print(f"rank {rank} it {it}: all-gather {t_ag:.3f} ms ({total / 1e6:.0f} MB, identical {same}); set from gathered "
          f"{e2.elapsed_time(e3):.3f} ms, from local {e4.elapsed_time(e5):.3f} ms (wall {wall:.3f} ms), ids {len(s1)} {len(s2)}",
          flush=True)
    s1.free(); s2.free(); s3.free()
if world > 1:
    dist.destroy_process_group()
