"""Device times of every BASELINE.json config's stages on one GPU (inputs resident in HBM, CUDA events on the
context's stream, best of 3): python tools/config_times.py [--scale 1.0]

  C1 alignment : 1 M pairs + PAF (-l 50 -c 0.5 -q 50)       set build, clean R1+R2
  C2 classifier: 10 M pairs + Kraken2 lines                  set build, clean (deplete, extract)
  C3 ONT       : long reads + many-line PAF                  set build (segmented OR stress), clean
  C4 set build : TXT id list (50 M ids at scale 1)           set build only (the 66 GB clean is C2 x 10)
  C5 diff      : input vs depleted output of C2              diff (ids of the output -> set, probe the input)
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from bench import taxids_for_config
from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=float, default=1.0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
ctx = api.Context(0)
res = {}


def timed(fn, reps=5):
    best, out = None, None
    for _ in range(reps):
        if out is not None and hasattr(out, "free"):
            out.free()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


def clean_all(ids, files, reverse=False):
    outs = [torch.empty(f.numel() + 64, dtype=torch.uint8, device=dev) for f in files]
    def go():
        return [api.clean_fastq_dev(ctx, ids, f, f.numel(), o, None, reverse) for f, o in zip(files, outs)]
    ms, rs = timed(go)
    return ms, rs, outs


def rec(name, ms, nbytes, **kw):
    res[name] = dict(ms=round(ms, 3), gb_per_s=round(nbytes / ms / 1e6, 1), **kw)
    print(name, res[name], flush=True)


# ---- clocks up before the sub-millisecond stages are timed
import time

_w = synth.gen_paf(1_000_000, device=dev)
_t0 = time.perf_counter()
while time.perf_counter() - _t0 < 1.5:  # by time: a GPU that idled at 120 MHz needs a moment under load
    api.IdSet.from_paf(ctx, _w, 50, 0.5, 50).free()
torch.cuda.synchronize()
del _w
# ---- C1
n1 = int(1_000_000 * a.scale)
fq = [synth.gen_fastq(n1, m, device=dev) for m in (1, 2)]
paf = synth.gen_paf(n1, device=dev)
ms, ids = timed(lambda: api.IdSet.from_paf(ctx, paf, 50, 0.5, 50))
rec("C1 set from PAF", ms, paf.numel(), lines=int((paf == 10).sum()), ids=len(ids))
ms, rs, _ = clean_all(ids, fq)
rec("C1 clean R1+R2", ms, sum(f.numel() for f in fq), reads=sum(r.reads_in for r in rs), kept=sum(r.reads_out for r in rs),
    path=rs[0].path)
ids.free()
del fq, paf
# ---- C2
n2 = int(10_000_000 * a.scale)
fq = [synth.gen_fastq(n2, m, device=dev) for m in (1, 2)]
kr = synth.gen_kraken_reads(n2, device=dev)
tax = taxids_for_config()
ms, ids = timed(lambda: api.IdSet.from_reads(ctx, kr, 0, tax))
rec("C2 set from Kraken2 lines", ms, kr.numel(), lines=n2, ids=len(ids))
ms, rs, outs = clean_all(ids, fq)
rec("C2 clean deplete", ms, sum(f.numel() for f in fq), reads=2 * n2, kept=sum(r.reads_out for r in rs), path=rs[0].path)
nw = [r.n_written for r in rs]
# ---- C5 (uses C2's depleted output)
pairs = [(fq[i], outs[i][: nw[i]]) for i in range(2)]
ms, d = timed(lambda: api.diff(ctx, pairs)[3])
rec("C5 diff (2 file pairs)", ms, sum(f.numel() for f in fq) + sum(nw), reads_in=2 * n2)
d.free()
ms, rs, _ = clean_all(ids, fq, True)
rec("C2 clean extract (-e)", ms, sum(f.numel() for f in fq), reads=2 * n2, kept=sum(r.reads_out for r in rs), path=rs[0].path)
ids.free()
del fq, kr, outs, pairs
torch.cuda.empty_cache()
# ---- C3
n3 = int(400_000 * a.scale)
ont, lens, uu = synth.gen_ont_fastq(n3, device=dev)
idl = [bytes(uu[i].tolist()) for i in range(0, n3, 2)]
ids = api.IdSet.from_ids(ctx, idl)
ms, rs, _ = clean_all(ids, [ont])
rec("C3 clean ONT", ms, ont.numel(), reads=n3, kept=rs[0].reads_out, path=rs[0].path, mean_len=int(lens.float().mean()))
ids.free()
del ont
# ---- C4 set build from a TXT id list
n4 = int(50_000_000 * a.scale)
txt = synth.gen_txt_ids(n4, device=dev)
ms, ids = timed(lambda: api.IdSet.from_txt(ctx, txt))
rec("C4 set from TXT ids", ms, txt.numel(), ids=len(ids))
ids.free()
print(json.dumps(res))
