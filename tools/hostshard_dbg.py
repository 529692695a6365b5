"""Emulates ONE rank of an N-rank C4 run on a single GPU: the rank's byte range of mate 1 through the host-buffer shard
entry point with a speculated line phase (SGPU_DEBUG=1 prints the chunk pipeline).  python tools/hostshard_dbg.py [world] [rank] [pairs]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import gen_fastq_range, gen_txt_full
from scrubby_b200 import api, synth
from scrubby_b200 import dist as sdist

world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rank = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000_000
dev = torch.device("cuda", 0)
ctx = api.Context(0)
ids = api.IdSet.from_txt(ctx, gen_txt_full(pairs, dev))
size = synth.fastq_size(pairs)
sh = sdist.plan_shards(size, world, halo=1 << 20)[rank]
d = gen_fastq_range(pairs, 1, sh.start, sh.start + sh.buf_len, dev)
print("shard", sh, "first bytes", bytes(d[:80].cpu().numpy()))
h = torch.empty(sh.buf_len + 16, dtype=torch.uint8, pin_memory=True)
h[: sh.buf_len].copy_(d[: sh.buf_len])
cap = int(sh.buf_len * 0.56) + (1 << 20)
h_out = torch.empty(cap, dtype=torch.uint8, pin_memory=True)
d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
rd = api.clean_fastq_shard_dev(ctx, ids, d, sh.buf_len, sh.own_len, 0 if sh.is_first else None, sh.is_first, sh.is_last, None, d_out)
print("device:", rd)
rh = api.clean_fastq_shard_host(ctx, ids, h, sh.buf_len, sh.own_len, 0 if sh.is_first else None, sh.is_first, sh.is_last, None, h_out)
print("host  :", rh)
