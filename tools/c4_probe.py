"""C4-like probe pressure at a fraction of C4's FASTQ size: the id set of `--set-pairs` pairs (default 100 M -> 50 M ids,
far larger than L2) against the first `--pairs` pairs of the mate files.  Prints the set build time and the fused
kernel's time per launch (CUDA events on the stream) for the library variant named by SGPU_VARIANT."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=10_000_000)
ap.add_argument("--set-pairs", type=int, default=100_000_000)
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--split", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = api.Context(0)
CH = 25_000_000
txt = torch.cat([synth.gen_txt_ids(min(CH, a.set_pairs - s), device=dev, start=s) for s in range(0, a.set_pairs, CH)])
d_r = [synth.gen_fastq(a.pairs, m, device=dev) for m in (1, 2)]
d_out = [torch.empty(int(t.numel() * 0.56) + (1 << 20), dtype=torch.uint8, device=dev) for t in d_r]
d_oth = [torch.empty(int(t.numel() * 0.56) + (1 << 20), dtype=torch.uint8, device=dev) if a.split else None for t in d_r]
torch.cuda.synchronize()
ctx.set_profiling(True)
torch.cuda.profiler.start()  # ncu --profile-from-start off: only the steps are captured
best_set, best_f = 1e9, 1e9
for s in range(a.steps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids = api.IdSet.from_txt(ctx, txt)
    e1.record()
    torch.cuda.synchronize()
    ctx.fused_stats()
    rs = [api.clean_fastq_dev(ctx, ids, d_r[i], d_r[i].numel(), d_out[i], d_oth[i]) for i in range(2)]
    f_ms, f_n, f_bytes = ctx.fused_stats()
    img = ids.image()
    ids.free()
    best_set = min(best_set, e0.elapsed_time(e1))
    best_f = min(best_f, f_ms / max(f_n, 1))
print(f"variant '{os.environ.get('SGPU_VARIANT', '')}': set build {best_set:.3f} ms ({img.count} ids, table {img.table_bytes / 1e9:.2f} GB), "
      f"fused {best_f:.3f} ms per launch = {f_bytes / f_n / best_f / 1e6:.1f} GB/s algorithmic, path {rs[0].path}, "
      f"kept {sum(r.reads_out for r in rs)} of {sum(r.reads_in for r in rs)}")
