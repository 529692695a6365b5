"""Device time of the GENERAL (always exact) path beside the fused kernel on the same canonical C2-shaped input, and
on the same records with "+id" separator lines (not a byte partition of the input: only the general path applies).

    python tools/general_time.py [--pairs 5000000]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import taxids_for_config
from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=5_000_000)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = api.Context(0)
fq = synth.gen_fastq(a.pairs, 1, device=dev)
ids = api.IdSet.from_reads(ctx, synth.gen_kraken_reads(a.pairs, device=dev), 0, taxids_for_config())
out = torch.empty(fq.numel() + (fq.numel() >> 3) + 64, dtype=torch.uint8, device=dev)  # room for the "+id" variant


def timed(label, buf, mode):
    ctx.set_mode(mode)
    best = None
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = api.clean_fastq_dev(ctx, ids, buf, buf.numel(), out, None)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    print(f"{label}: {best:.3f} ms, {buf.numel() / best / 1e6:.1f} GB/s in, {a.pairs / best / 1e3:.1f} M reads/s, path {r.path}, "
          f"kept {r.reads_out}, written {r.n_written}", flush=True)
    return r


r1 = timed("canonical, fused (auto)", fq, 0)
r2 = timed("canonical, general (forced)", fq, 1)
assert (r1.reads_in, r1.reads_out, r1.n_written) == (r2.reads_in, r2.reads_out, r2.n_written)
# "+id" separator lines: the 8-digit class only, so that the records have one length
n = min(a.pairs, 5_000_000)
fq8 = synth.gen_fastq(n, 1, device=dev, start=10_000_000)
L = fq8.numel() // n
rows = fq8.view(n, L)
d = 8
sep = torch.cat([rows[:, : 19 + d + 152], rows[:, 1: 5 + d], rows[:, 19 + d + 152:]], dim=1).contiguous().view(-1)
r3 = timed('"+id" separators, auto (falls back to general)', sep, 0)
assert r3.path == 2 and r3.reads_in == n
