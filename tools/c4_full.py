"""Configs 4 and 5 of BASELINE.json at FULL size on ONE GPU: 100 M 2x150 pairs (2 x 33.1 GB of FASTQ) against a
50 M-id depletion set, then `diff` of every mate file against its depleted output.  Inputs are generated in HBM;
times are CUDA events on the context's stream.  Results are checked through size-independent properties (the CPU
oracle would need ~10 minutes per file at this size):

  * reads_in == N, reads_out == number of reads that are not host (recomputed here from the membership rule),
  * bytes written == the sum of the record sizes of the kept reads (closed form per id-length class),
  * the output holds exactly 4 * reads_out newlines and is a fixed point (cleaning it again returns it unchanged),
  * diff: reads_in / reads_out / difference == N / kept / N - kept, and |absent ids| == |set|.

    python tools/c4_full.py [--pairs 100000000]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from scrubby_b200 import api, synth

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=100_000_000)
a = ap.parse_args()
N = a.pairs
dev = torch.device("cuda", 0)
ctx = api.Context(0)
res = {"pairs": N}


def ev():
    return torch.cuda.Event(enable_timing=True)


def timed(fn, reps=1, free=None):
    """best of `reps` (the first call of a size also grows the library's memory pool)"""
    best, out = None, None
    for _ in range(reps):
        if out is not None and free is not None:
            free(out)
        e0, e1 = ev(), ev()
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return best, out


# ---- the depletion set (config 4: "50M-ID depletion set"), delivered as a TXT id list
CH = 25_000_000
txt = torch.cat([synth.gen_txt_ids(min(CH, N - s), device=dev, start=s) for s in range(0, N, CH)])
api.IdSet.from_txt(ctx, txt[: 1 << 20]).free()  # warm-up
ms, ids = timed(lambda: api.IdSet.from_txt(ctx, txt), 3, lambda o: o.free())
res["set_from_txt"] = dict(ms=round(ms, 3), ids=len(ids), txt_bytes=txt.numel())
print("C4 set from TXT", res["set_from_txt"], flush=True)
del txt

# ---- expected counts from the membership rule
kept_expected, bytes_expected = 0, 0
for s in range(0, N, CH):
    idx = torch.arange(s, min(N, s + CH), dtype=torch.int64, device=dev)
    keep = ~synth.is_host(idx)
    kept_expected += int(keep.sum())
    digits = torch.ones_like(idx)
    for k in range(1, 10):
        digits += (idx >= 10 ** k).long()
    bytes_expected += int(((323 + digits) * keep).sum())
    del idx, keep, digits
assert len(ids) == N - kept_expected, (len(ids), N - kept_expected)

size = synth.fastq_size(N)
d_in = torch.empty(size + 64, dtype=torch.uint8, device=dev)
d_out = torch.empty(size + 64, dtype=torch.uint8, device=dev)
d_again = torch.empty(bytes_expected + 64, dtype=torch.uint8, device=dev)
tot_ms, tot_diff_ms = 0.0, 0.0
for mate in (1, 2):
    fq = synth.gen_fastq(N, mate, device=dev, out=d_in)
    torch.cuda.synchronize()
    if mate == 1:  # warm-up on a prefix that ends on a record boundary
        pre = synth.fastq_size(1_000_000)
        api.clean_fastq_dev(ctx, ids, fq[:pre], pre, d_out, None)
    ms, r = timed(lambda: api.clean_fastq_dev(ctx, ids, fq, fq.numel(), d_out, None), 2)
    tot_ms += ms
    assert r.path == 1, "the fused kernel must take canonical input"
    assert (r.reads_in, r.reads_out, r.n_written) == (N, kept_expected, bytes_expected), (r.reads_in, r.reads_out, r.n_written)
    out = d_out[: r.n_written]
    nl = sum(int((out[o: o + (1 << 30)] == 10).sum()) for o in range(0, out.numel(), 1 << 30))
    assert nl == 4 * kept_expected, nl
    r2 = api.clean_fastq_dev(ctx, ids, out, out.numel(), d_again, None)
    assert (r2.reads_in, r2.reads_out, r2.n_written) == (kept_expected, kept_expected, bytes_expected)
    assert all(torch.equal(out[o: o + (1 << 30)], d_again[o: min(o + (1 << 30), out.numel())])
               for o in range(0, out.numel(), 1 << 30)), "the depleted output is not a fixed point"
    res[f"clean_R{mate}"] = dict(ms=round(ms, 3), reads=N, kept=r.reads_out, in_bytes=fq.numel(), out_bytes=r.n_written,
                                 gb_per_s_in=round(fq.numel() / ms / 1e6, 1),
                                 gb_per_s_alg=round((fq.numel() + r.n_written) / ms / 1e6, 1))
    print(f"C4 clean R{mate}", res[f"clean_R{mate}"], flush=True)
    # ---- config 5: diff of this mate file against its depleted output
    ms, d = timed(lambda: api.diff(ctx, [(fq, out)]), 2, lambda o: o[3].free())
    tot_diff_ms += ms
    assert d[:3] == (N, kept_expected, N - kept_expected), d[:3]
    assert len(d[3]) == N - kept_expected
    d[3].free()
    res[f"diff_R{mate}"] = dict(ms=round(ms, 3), reads_in=N, reads_out=kept_expected, difference=N - kept_expected,
                                gb_per_s=round((fq.numel() + out.numel()) / ms / 1e6, 1))
    print(f"C5 diff R{mate}", res[f"diff_R{mate}"], flush=True)
res["clean_total"] = dict(ms=round(tot_ms, 3), reads_per_s=round(2 * N / tot_ms * 1e3), fastq_gb_per_s=round(2 * size / tot_ms / 1e6, 1))
res["diff_total"] = dict(ms=round(tot_diff_ms, 3), reads_per_s=round(2 * N / tot_diff_ms * 1e3))
res["torch_peak_allocated_gb"] = round(torch.cuda.max_memory_allocated() / 1e9, 1)
print("C4+C5 full size, one GPU:", json.dumps(res))
