"""Minimal driver for ncu: the bench step (Kraken2 lines -> id set -> clean R1, R2) on device-resident data.

    ncu ... python tools/prof_step.py --pairs 2000000 --steps 2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import taxids_for_config
from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=2_000_000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--split", action="store_true")
ap.add_argument("--ont", type=int, default=0, help="use N ONT-like long reads instead of 2x150 pairs")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = api.Context(0)
taxids = taxids_for_config()
if a.ont:
    fq, lens, uu = synth.gen_ont_fastq(a.ont, device=dev)
    d_r = [fq]
    ids_list = [bytes(uu[i].tolist()) for i in range(0, a.ont, 2)]
    mk = lambda: api.IdSet.from_ids(ctx, ids_list)
else:
    d_r = [synth.gen_fastq(a.pairs, m, device=dev) for m in (1, 2)]
    d_k = synth.gen_kraken_reads(a.pairs, device=dev)
    mk = lambda: api.IdSet.from_reads(ctx, d_k, 0, taxids)
d_out = [torch.empty(t.numel() + 64, dtype=torch.uint8, device=dev) for t in d_r]
d_oth = [torch.empty(t.numel() + 64, dtype=torch.uint8, device=dev) if a.split else None for t in d_r]
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the steps are captured
ctx.set_profiling(True)
ctx.fused_stats()
for s in range(a.steps):
    ids = mk()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rs = [api.clean_fastq_dev(ctx, ids, d_r[i], d_r[i].numel(), d_out[i], d_oth[i]) for i in range(len(d_r))]
    e1.record()
    torch.cuda.synchronize()
    ids.free()
    nbytes = sum(t.numel() for t in d_r)
    f_ms, f_n, f_bytes = ctx.fused_stats()
    if f_n:
        print(f"step {s}: fused kernel {f_n} launches, avg {f_ms / f_n:.3f} ms, {f_bytes / f_ms / 1e6:.1f} GB/s algorithmic")
    print(f"step {s}: clean {e0.elapsed_time(e1):.3f} ms, {nbytes / e0.elapsed_time(e1) / 1e6:.1f} GB/s in, path {rs[0].path}, "
          f"reads {sum(r.reads_in for r in rs)} kept {sum(r.reads_out for r in rs)}")
torch.cuda.profiler.stop()
