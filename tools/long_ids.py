"""Real-shape read names: 2x150 records with 40-byte Illumina-style ids -- the fingerprint + key-arena
path of the id set -- beside the inline-id (<= 15 bytes) records of the bench, same record count, same set size.
Prints the fused kernel's time per launch and its fraction of the measured HBM peak for both.
    python tools/long_ids.py [--pairs 10000000]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import peaks
from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = api.Context(0)
peak, _ = peaks()
res = {}
for name, gen_fq, gen_ids in (("inline ids (syn.N, <= 13 bytes)", synth.gen_fastq, synth.gen_txt_ids),
                              ("Illumina names (40 bytes)", synth.gen_fastq_illumina, synth.gen_txt_ids_illumina)):
    txt = gen_ids(a.pairs, device=dev)
    fq = gen_fq(a.pairs, 1, device=dev)
    n = int(fq.numel())
    pad = torch.zeros(n + 16, dtype=torch.uint8, device=dev)
    pad[:n] = fq
    del fq
    out = torch.empty(int(n * 0.56) + (1 << 20), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ctx.set_profiling(True)
    best_f, best_s = 1e9, 1e9
    for _ in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids = api.IdSet.from_txt(ctx, txt)
        e1.record()
        torch.cuda.synchronize()
        ctx.fused_stats()
        r = api.clean_fastq_dev(ctx, ids, pad, n, out, None)
        f_ms, f_n, f_bytes = ctx.fused_stats()
        img = ids.image()
        ids.free()
        best_f, best_s = min(best_f, f_ms / max(f_n, 1)), min(best_s, e0.elapsed_time(e1))
    assert r.path == 1 and r.reads_in == a.pairs
    gbs = f_bytes / f_n / best_f / 1e6
    res[name] = dict(records=a.pairs, fastq_gb=round(n / 1e9, 2), kept=r.reads_out, fused_ms=round(best_f, 3),
                     algorithmic_gb_per_s=round(gbs, 1), frac_of_measured_hbm=round(gbs / peak, 3), set_build_ms=round(best_s, 3),
                     table_gb=round(img.table_bytes / 1e9, 2), arena_gb=round(img.arena_bytes / 1e9, 2))
    print(name, res[name], flush=True)
    del pad, out, txt
print(json.dumps(res))
