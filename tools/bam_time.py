"""Time of sgpu_idset_from_bam (host buffer: host walk of the block_size chain + H2D + bam_parse_kernel + set build) on a
synthetic BAM stream of minimap2-sr-like records, and of the host stage in front of it (parallel BGZF inflate).

    python tools/bam_time.py [--records 2000000]
"""
import argparse
import os
import struct
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

import bam_build as bb
from scrubby_b200 import api, hostlib

ap = argparse.ArgumentParser()
ap.add_argument("--records", type=int, default=2_000_000)
a = ap.parse_args()
n = a.records
tmpl = np.frombuffer(bb.record(b"syn.00000000", cigar="20S100M5I25S", aux=b"NMi\x03\x00\x00\x00ASC\x60"), dtype=np.uint8)
recs = np.tile(tmpl, (n, 1))
idx = np.arange(n, dtype=np.int64)
for k in range(8):  # the digits of the read name (offset 4 block_size + 32 fixed + 4 "syn.")
    recs[:, 40 + k] = (idx // 10 ** (7 - k)) % 10 + 48
recs[:, 4 + 9] = np.where(idx % 7 == 0, 3, 60)                 # mapq
recs[:, 4 + 14] = np.where(idx % 10 == 0, 4, 0).astype(np.uint8)  # flag: every tenth record unmapped
stream = bb.stream([], refs=((b"chr1", 248956422),)) + recs.tobytes()
ctx = api.Context(0)
h = torch.frombuffer(bytearray(stream), dtype=torch.uint8).pin_memory()
api.IdSet.from_bam(ctx, h[: len(bb.stream([], refs=((b"chr1", 248956422),))) + 100 * tmpl.size], 50, 0.5, 50).free()
best = None
for _ in range(4):
    t0 = time.perf_counter()
    s = api.IdSet.from_bam(ctx, h, 50, 0.5, 50)
    dt = time.perf_counter() - t0
    best = dt if best is None else min(best, dt)
    k = len(s)
    s.free()
expect = int(((idx % 7 != 0) & (idx % 10 != 0)).sum())
assert k == expect, (k, expect)
print(f"from_bam: {n} records, {len(stream) / 1e6:.0f} MB decompressed, {best * 1e3:.2f} ms end to end from pinned host memory "
      f"= {len(stream) / best / 1e9:.1f} GB/s, {n / best / 1e6:.1f} M records/s, {k} ids")
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "x.bam")
    part = stream[: 200 * 1000 * 1000]
    with open(p, "wb") as f:
        f.write(bb.bgzf(part))
    t0 = time.perf_counter()
    out = hostlib.read_file(p)
    dt = time.perf_counter() - t0
    assert out == part
    print(f"host stage: BGZF inflate of {len(part) / 1e6:.0f} MB on {os.cpu_count()} host threads: {dt * 1e3:.0f} ms = {len(part) / dt / 1e9:.2f} GB/s")
