"""Small inputs through every kernel family for compute-sanitizer (memcheck / racecheck / synccheck / initcheck):
the fused kernel (kept, kept+removed, ids mode, a shard with a speculated phase, the chunked host pipeline), the general
path, evidence parsers, the id set (key-by-key CAS insert, bulk page build, rehash, dump).  Results are checked against
the oracle so that a sanitizer run is also a parity run.
    compute-sanitizer --tool racecheck python tools/sanitize_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("SGPU_PIPE_CHUNK", "65536")
os.environ.setdefault("SGPU_PIPE_HALO", "8192")
import torch

from oracle import oracle as orc
from scrubby_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
ctx = api.Context(0)
fq = [synth.gen_fastq(n, m).numpy().tobytes() for m in (1, 2)]
txt = synth.gen_txt_ids(n).numpy().tobytes()
paf = synth.gen_paf(n).numpy().tobytes()
kr = synth.gen_kraken_reads(n).numpy().tobytes()
oset = orc.set_from_txt(txt)
# id set: bulk build (SGPU_IDSET_BULK_MIN=1 in the environment) or CAS insert, long ids, dump
gs = api.IdSet.from_txt(ctx, txt)
assert gs.sorted_ids() == oset.sorted_ids()
gp = api.IdSet.from_paf(ctx, paf, 50, 0.5, 50)
assert gp.sorted_ids() == orc.set_from_paf(paf, 50, 0.5, 50).sorted_ids()
ill = synth.gen_txt_ids_illumina(n).numpy().tobytes()
gi = api.IdSet.from_txt(ctx, ill)
assert gi.sorted_ids() == orc.set_from_txt(ill).sorted_ids()
# fused kernel: both outputs, extract mode, ids mode (diff), general path
for reverse in (False, True):
    g = api.clean_fastq(ctx, gs, fq[0], reverse)
    o = orc.clean_fastq(fq[0], oset, reverse)
    assert g.written == o.written and g.other == o.other and g.path == 1
fqi = synth.gen_fastq_illumina(n, 1).numpy().tobytes()
g = api.clean_fastq(ctx, gi, fqi)
assert g.written == orc.clean_fastq(fqi, orc.set_from_txt(ill)).written and g.path == 1
ctx.set_mode(1)
g = api.clean_fastq(ctx, gs, fq[1])
assert g.written == orc.clean_fastq(fq[1], oset).written and g.path == 2
ctx.set_mode(0)
pairs = [(fq[0], orc.clean_fastq(fq[0], oset).written)]
d = api.diff(ctx, pairs)
od = orc.diff(pairs)
assert d[:3] == od[:3] and d[3].sorted_ids() == od[3].sorted_ids()
# shards: speculated phase on device buffers and through the chunked host pipeline
whole = orc.clean_fastq(fq[0], oset).written
cut = (len(fq[0]) // 2) & ~15
got = b""
for host in (False, True):
    got = b""
    for s, (a, b) in enumerate([(0, cut), (cut, len(fq[0]))]):
        end = len(fq[0]) if s else b + 4096
        t = torch.zeros(end - a + 16, dtype=torch.uint8)
        t[: end - a] = torch.frombuffer(bytearray(fq[0][a:end]), dtype=torch.uint8)
        if host:
            h_in, h_out = t.pin_memory(), torch.empty(end - a + 64, dtype=torch.uint8).pin_memory()
            r = api.clean_fastq_shard_host(ctx, gs, h_in, end - a, b - a, 0 if s == 0 else None, s == 0, s == 1, None, h_out)
            got += h_out[: r.n_written].numpy().tobytes()
        else:
            d_in, d_out = t.cuda(), torch.empty(end - a + 64, dtype=torch.uint8, device="cuda")
            r = api.clean_fastq_shard_dev(ctx, gs, d_in, end - a, b - a, 0 if s == 0 else None, s == 0, s == 1, None, d_out)
            got += d_out[: r.n_written].cpu().numpy().tobytes()
        assert r.status == 0 and r.path == 1
    assert got == whole
# Kraken2 lines
from scrubby_b200 import hostlib

rep = synth.gen_kraken_report(500)
tax = hostlib.get_taxids_from_report(rep, ["Chordata"], ["9606"])
gk = api.IdSet.from_reads(ctx, kr, 0, tax)
assert gk.sorted_ids() == orc.set_from_reads(kr, 0, orc.taxids_from_report(rep, ["Chordata"], ["9606"])).sorted_ids()
print("sanitize_step ok:", n, "records, launches", ctx.launches)
