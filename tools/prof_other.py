"""Driver for the ncu capture of the kernels beside the fused one: Kraken2 lines -> set (tile parser, bulk set build),
a key-by-key insert (SGPU path for small / growing sets), and the general FASTQ path (record + copy kernels)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import taxids_for_config
from scrubby_b200 import api, synth

dev = torch.device("cuda", 0)
ctx = api.Context(0)
n = 5_000_000
kr = synth.gen_kraken_reads(n, device=dev)
fq = synth.gen_fastq(n, 1, device=dev)
out = torch.empty(fq.numel() + 64, dtype=torch.uint8, device=dev)
tax = taxids_for_config()
ids = api.IdSet.from_reads(ctx, kr, 0, tax)  # warm-up
ids.free()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ids = api.IdSet.from_reads(ctx, kr, 0, tax)          # lines_tile_kernel, idset_count / scatter / page
os.environ["SGPU_IDSET_BULK_MIN"] = "0"
more = api.IdSet.from_ids(ctx, [b"extra.%d" % i for i in range(200000)])
ctx.set_mode(1)
r = api.clean_fastq_dev(ctx, ids, fq, fq.numel(), out, None)  # fastq_record_kernel, fastq_copy_kernel
ctx.set_mode(0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("prof_other ok", len(ids), r.path, r.reads_out)
