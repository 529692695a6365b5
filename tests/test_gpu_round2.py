"""Round-2 GPU tests (through the C ABI): the dense 128-byte-bucket id set, output capacity handling, the one-pass
(speculated line phase) shard protocol on device and host buffers, and the oracle-vs-GPU parity of BASELINE's five
configs AT FULL SIZE (VERDICT r01, row g): the C oracle runs over the same generator output in record-aligned chunks on
several host threads and every chunk's bytes are compared with the GPU's output on the device.
"""
import os
import random
import sys
from concurrent.futures import ThreadPoolExecutor

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu

from oracle import oracle as orc  # noqa: E402  (test infrastructure: the checker)
from scrubby_b200 import _lib, api, synth  # noqa: E402
from scrubby_b200 import dist as sdist  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    c = api.Context(0)
    yield c
    c.close()


def _dev(b: bytes, pad: int = 16):
    t = torch.zeros(len(b) + pad, dtype=torch.uint8, device="cuda")
    if b:
        t[: len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
    return t


# ------------------------------------------------------------------------------------------------ dense id set
def test_dense_idset_membership_and_dump(ctx):
    """200 k keys of every length class (1..15 inline, 16..60 through the arena), duplicates, near-collisions: the set's
    sorted dump and membership equal the oracle's HashSet; the table is sized exactly for load 0.2 in pages of 128-byte buckets"""
    rng = random.Random(11)
    keys = set()
    for i in range(120_000):
        keys.add(b"syn.%d" % i)
    for i in range(40_000):
        n = rng.randrange(1, 61)
        keys.add(bytes(rng.randrange(33, 127) for _ in range(n)))
    for i in range(40_000):  # long ids that share a 40-byte prefix (only the arena compare tells them apart)
        keys.add(b"M01234:77:000000000-ABCDE:1:1101:15589:%05d" % i)
    lst = list(keys)
    rng.shuffle(lst)
    lst += lst[:5000]  # duplicates
    g = api.IdSet.from_ids(ctx, lst)
    o = orc.OSet.from_ids(lst)
    assert len(g) == len(o) == len(keys)
    assert g.sorted_ids() == o.sorted_ids()
    for k in rng.sample(lst, 200):
        assert k in g
    for k in (b"syn.120000", b"syn", b"M01234:77:000000000-ABCDE:1:1101:15589:99999", b"x" * 70):
        assert (k in g) == (k in o)
    img = g.image()
    assert img.capacity % 8 == 0 and img.table_bytes == img.capacity * 16
    assert 0.15 <= len(keys) / img.capacity <= 0.2 + 1e-9, "load factor of an exactly sized table"
    g2 = api.IdSet.from_image(ctx, img)
    assert g2.sorted_ids() == o.sorted_ids()


def test_dense_idset_grows_across_inserts(ctx):
    """diff accumulates ids of several file pairs into one set: the table is rehashed as it grows"""
    n = 30_000
    ids_all = orc.OSet()
    h = None
    pairs = []
    for k in range(3):
        fq = synth.gen_fastq(n, 1, start=k * n).numpy().tobytes()
        sel = orc.set_from_txt(synth.gen_txt_ids(n, start=k * n).numpy().tobytes())
        pairs.append((fq, orc.clean_fastq(fq, sel).written))
    want = orc.diff(pairs)
    got = api.diff(ctx, pairs)
    assert got[:3] == want[:3]
    assert got[3].sorted_ids() == want[3].sorted_ids()
    del ids_all, h


# ------------------------------------------------------------------------------------------------ output capacity
def test_output_smaller_than_input_is_enough(ctx):
    """the single-pass kernel no longer needs cap >= n_in: a buffer that just holds the kept bytes is accepted (path 1),
    one byte less is SGPU_ERR_CAPACITY and nothing is written past the buffer"""
    n = 40_000
    fq = synth.gen_fastq(n, 1, device="cuda")
    gs = api.IdSet.from_txt(ctx, synth.gen_txt_ids(n, device="cuda"))
    big = torch.empty(fq.numel() + 64, dtype=torch.uint8, device="cuda")
    r0 = api.clean_fastq_dev(ctx, gs, fq, fq.numel(), big, None)
    assert r0.path == 1 and r0.n_written < fq.numel() * 0.6
    exact = torch.empty(r0.n_written + 32, dtype=torch.uint8, device="cuda")
    exact[r0.n_written:] = 0xAB
    r1 = api.clean_fastq_dev(ctx, gs, fq, fq.numel(), exact[: r0.n_written], None)
    assert r1.path == 1 and r1.n_written == r0.n_written
    assert torch.equal(exact[: r0.n_written], big[: r0.n_written])
    assert bool((exact[r0.n_written:] == 0xAB).all()), "bytes behind the buffer were touched"
    small = torch.full((r0.n_written + 32,), 0xCD, dtype=torch.uint8, device="cuda")
    with pytest.raises(api.ScrubbyGpuError) as e:
        api.clean_fastq_dev(ctx, gs, fq, fq.numel(), small[: r0.n_written - 1], None)
    assert e.value.status == _lib.SGPU_ERR_CAPACITY
    assert bool((small[r0.n_written - 1:] == 0xCD).all()), "bytes behind the buffer were touched"


# ------------------------------------------------------------------------------------------------ one-pass shards
def _expected_spec(buf: bytes):
    """the speculation rule restated: (record start, newlines before it) or None"""
    pos, p = [], -1
    for _ in range(4):
        p = buf.find(b"\n", p + 1)
        if p < 0:
            break
        pos.append(p)
    j = next((i for i, q in enumerate(pos) if buf[q + 1: q + 3] == b"+\n"), None)
    if j is None:
        return None
    k = (2 + j) & 3
    return (pos[k] + 1, k + 1) if k < len(pos) else None


def _one_pass(ctx, gs, fq: bytes, bounds, halo, host=False, want_other=True):
    """runs every shard with SGPU_NEWLINES_UNKNOWN / crlf = -1 and applies the caller's check; returns
    (accepted, written, other, reads_in, reads_out)"""
    k = len(bounds) - 1
    outs, others, rin, rout = [], [], 0, 0
    before, ok = 0, True
    for s in range(k):
        a, b = bounds[s], bounds[s + 1]
        end = len(fq) if s == k - 1 else min(len(fq), b + halo)
        cap = end - a + 64
        if host:
            h_in = torch.zeros(end - a + 16, dtype=torch.uint8).pin_memory()
            h_in[: end - a] = torch.frombuffer(bytearray(fq[a:end]), dtype=torch.uint8)
            h_out = torch.empty(cap, dtype=torch.uint8).pin_memory()
            h_oth = torch.empty(cap, dtype=torch.uint8).pin_memory() if want_other else None
            r = api.clean_fastq_shard_host(ctx, gs, h_in, end - a, b - a, 0 if s == 0 else None, s == 0, s == k - 1, None,
                                           h_out, h_oth)
            w, o = h_out, h_oth
        else:
            d_in = _dev(fq[a:end])
            d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
            d_oth = torch.empty(cap, dtype=torch.uint8, device="cuda") if want_other else None
            r = api.clean_fastq_shard_dev(ctx, gs, d_in, end - a, b - a, 0 if s == 0 else None, s == 0, s == k - 1, None,
                                          d_out, d_oth)
            w, o = d_out, d_oth
        if r.status == _lib.SGPU_ERR_PHASE_UNKNOWN or r.path != 1:
            return False, None, None, 0, 0  # declined: the exact protocol takes over
        assert r.status == 0
        assert r.own_newlines == fq[a:b].count(b"\n"), "own-range newline count"
        if s == 0:
            assert not r.speculated and r.lead_newlines == 0 and not r.crlf
        else:
            assert r.speculated
            exp = _expected_spec(fq[a:end])
            assert exp is not None and r.lead_newlines == exp[1]
            ok = ok and (before + r.lead_newlines) % 4 == 0
        before += r.own_newlines
        outs.append(w[: r.n_written].cpu().numpy().tobytes())
        others.append(o[: r.n_other].cpu().numpy().tobytes() if want_other else b"")
        rin += r.reads_in
        rout += r.reads_out
    return ok, b"".join(outs), b"".join(others), rin, rout


@pytest.mark.parametrize("cut_mode", ["random", "at_start", "after_start", "before_start", "in_plus", "in_qual"])
@pytest.mark.parametrize("host", [False, True])
def test_one_pass_shards_match_whole(ctx, cut_mode, host):
    """every cut position class: the speculated phase is accepted by the newline-count check and the concatenated
    shards equal the unsharded run (kept and removed streams, counters)"""
    n = 6000
    fq = synth.gen_fastq(n, 1, start=995).numpy().tobytes()
    ids = synth.gen_txt_ids(n + 995).numpy().tobytes()
    gs = api.IdSet.from_txt(ctx, ids)
    whole = api.clean_fastq(ctx, gs, fq)
    assert whole.written == orc.clean_fastq(fq, orc.set_from_txt(ids)).written
    rng = random.Random(hash(cut_mode) & 0xFFFF)
    starts = [0]
    p = 0
    while True:
        p = fq.find(b"\n@syn.", p)
        if p < 0:
            break
        starts.append(p + 1)
        p += 1
    k = 5
    cuts = []
    for _ in range(k - 1):
        s = starts[rng.randrange(1, len(starts))]
        if cut_mode == "random":
            c = rng.randrange(1, len(fq))
        elif cut_mode == "in_plus":
            c = fq.find(b"\n+\n", s) + rng.randrange(1, 3)  # on the '+' or on the newline behind it
        elif cut_mode == "in_qual":
            c = fq.find(b"\n+\n", s) + 3 + rng.randrange(0, 150)
        else:
            c = s + {"at_start": 0, "after_start": 1, "before_start": -1}[cut_mode]
        cuts.append(c)
    bounds = [0] + sorted(set(cuts)) + [len(fq)]
    # device shards must start on 16-byte boundaries of the FILE only when they are views of one buffer; here every
    # shard is uploaded on its own, so any cut is legal
    ok, w, o, rin, rout = _one_pass(ctx, gs, fq, bounds, halo=4096, host=host)
    assert ok, "a canonical file's speculation must be accepted"
    assert w == whole.written and o == whole.other
    assert (rin, rout) == (whole.reads_in, whole.reads_out)


def test_one_pass_refuted_or_declined_on_ambiguous_records(ctx):
    """one-base reads with quality "+": a shard that starts on a separator's '+' sees a QUALITY line first.  Such a file is
    never accepted with wrong bytes: the single-pass kernel declines it (lines this short are not canonical for it), or
    the check on the exchanged newline counts refutes the speculation"""
    # (a 110-byte comment keeps the records above the single-pass kernel's minimum of ~102 bytes per record)
    cm = b" " + b"c" * 110
    recs = [b"@r%d%s\nA\n+\n%s\n" % (i, cm, b"+" if i % 3 else b"I") for i in range(3000)]
    fq = b"".join(recs)
    ids = b"".join(b"r%d\n" % i for i in range(0, 3000, 2))
    gs = api.IdSet.from_txt(ctx, ids)
    whole = api.clean_fastq(ctx, gs, fq)
    assert whole.written == orc.clean_fastq(fq, orc.set_from_txt(ids)).written
    refuted = accepted = 0
    for rec in range(100, 2900, 97):
        s = fq.find(b"@r%d " % rec)
        h = len(b"@r%d" % rec) + len(cm)
        for cut in (s, s + h + 3, s + h + 4, s + h + 5):  # record start, the separator's '+', its newline, the quality
            ok, w, o, rin, rout = _one_pass(ctx, gs, fq, [0, cut, len(fq)], halo=4096)
            if ok:
                accepted += 1
                assert w == whole.written and o == whole.other and (rin, rout) == (whole.reads_in, whole.reads_out)
            else:
                refuted += 1
    assert refuted, "the ambiguous cuts must be declined or refuted"


def test_host_shard_pipeline_small_chunks(ctx, monkeypatch):
    """sgpu_clean_fastq_shard with the chunked H2D / kernel / D2H pipeline forced onto a small shard (16 KiB chunks):
    identical to the device shard call, speculated and exact"""
    monkeypatch.setenv("SGPU_PIPE_CHUNK", "16384")
    monkeypatch.setenv("SGPU_PIPE_HALO", "4096")
    n = 3000
    fq = synth.gen_fastq(n, 2, start=77).numpy().tobytes()
    ids = synth.gen_txt_ids(n + 77).numpy().tobytes()
    gs = api.IdSet.from_txt(ctx, ids)
    whole = api.clean_fastq(ctx, gs, fq)
    assert whole.written == orc.clean_fastq(fq, orc.set_from_txt(ids)).written
    cut = len(fq) // 3 + 5
    cut2 = 2 * len(fq) // 3 + 11
    ok, w, o, rin, rout = _one_pass(ctx, gs, fq, [0, cut, cut2, len(fq)], halo=2048, host=True)
    assert ok and w == whole.written and o == whole.other and (rin, rout) == (whole.reads_in, whole.reads_out)
    # exact phase through the same entry point
    outs = []
    nb = 0
    bounds = [0, cut, cut2, len(fq)]
    for s in range(3):
        a, b = bounds[s], bounds[s + 1]
        end = len(fq) if s == 2 else b + 2048
        h_in = torch.zeros(end - a + 16, dtype=torch.uint8).pin_memory()
        h_in[: end - a] = torch.frombuffer(bytearray(fq[a:end]), dtype=torch.uint8)
        h_out = torch.empty(end - a + 64, dtype=torch.uint8).pin_memory()
        r = api.clean_fastq_shard_host(ctx, gs, h_in, end - a, b - a, nb, s == 0, s == 2, False, h_out, None)
        assert r.status == 0
        nb += fq[a:b].count(b"\n")
        outs.append(h_out[: r.n_written].numpy().tobytes())
    assert b"".join(outs) == whole.written


def test_host_shard_pipeline_default_chunks(ctx):
    """the host shard entry point with its real 128 MiB chunks: a 1 GB file in two shards, each pipelined over several
    chunks; newline counts, speculation check and bytes as for the unsharded call"""
    n = int(os.environ.get("SGPU_TEST_PIPE_RECORDS", "3000000"))
    d_fq = synth.gen_fastq(n, 1, device="cuda")
    fq_t = d_fq.cpu()
    gs = api.IdSet.from_txt(ctx, synth.gen_txt_ids(n, device="cuda"))
    size = int(fq_t.numel())
    d_all = torch.zeros(size + 16, dtype=torch.uint8, device="cuda")
    d_all[:size] = d_fq
    del d_fq
    d_w = torch.empty(size + 64, dtype=torch.uint8, device="cuda")
    whole = api.clean_fastq_dev(ctx, gs, d_all, size, d_w, None)
    cut = (size // 2) & ~15
    halo = 1 << 20
    before, off = 0, 0
    for s, (a, b) in enumerate([(0, cut), (cut, size)]):
        end = size if s == 1 else b + halo
        h_in = torch.zeros(end - a + 16, dtype=torch.uint8, pin_memory=True)
        h_in[: end - a] = fq_t[a:end]
        h_out = torch.empty(int((end - a) * 0.6), dtype=torch.uint8, pin_memory=True)
        r = api.clean_fastq_shard_host(ctx, gs, h_in, end - a, b - a, 0 if s == 0 else None, s == 0, s == 1, None, h_out, None)
        assert r.status == 0 and r.path == 1
        assert r.own_newlines == int((fq_t[a:b] == 10).sum()), "own-range newline count over several chunks"
        if s:
            assert r.speculated and (before + r.lead_newlines) % 4 == 0
        before += r.own_newlines
        assert torch.equal(h_out[: r.n_written].cuda(), d_w[off: off + r.n_written])
        off += r.n_written
    assert off == whole.n_written


# ------------------------------------------------------------------------------------------------ full-size parity
N_THREADS = max(2, min(12, (os.cpu_count() or 4) - 2))


def _record_bounds(n: int, start: int, per: int):
    """byte offsets of every `per`-th record of the synthetic file syn.{start}.. (record-aligned chunk bounds)"""
    return [synth.fastq_size(min(n, i), start) for i in range(0, n + per, per)]


def _parity_full(ctx, gs, oset, n, mate, reverse, split, per=1_000_000):
    fq = synth.gen_fastq(n, mate, device="cuda")
    pad = torch.zeros(fq.numel() + 16, dtype=torch.uint8, device="cuda")
    pad[: fq.numel()] = fq
    del fq
    n_in = pad.numel() - 16
    d_w = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    d_o = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda") if split else None
    r = api.clean_fastq_dev(ctx, gs, pad, n_in, d_w, d_o, reverse)
    assert r.path == 1
    bounds = _record_bounds(n, 0, per)
    rin, rout, nw, no = _chunks(pad, bounds, oset, reverse, d_w, d_o)
    assert (rin, rout, nw) == (r.reads_in, r.reads_out, r.n_written)
    if split:
        assert no == r.n_other and nw + no == n_in
    return r


def _chunks(d_in, bounds, oset, reverse, d_w, d_o):
    """the oracle over record-aligned chunks [bounds[i], bounds[i+1]) of the device file `d_in` on N_THREADS host
    threads; every chunk's written / other bytes must equal the next bytes of the GPU outputs d_w / d_o.
    Returns (reads_in, reads_out, bytes_written, bytes_other)."""
    def one(i):
        a, b = bounds[i], bounds[i + 1]
        return orc.clean_fastq(d_in[a:b].cpu().numpy(), oset, reverse, want_bytes=False)

    tot = [0, 0, 0, 0]
    with ThreadPoolExecutor(N_THREADS) as ex:
        for k, rr in enumerate(ex.map(one, range(len(bounds) - 1))):  # map yields in order
            w = torch.from_numpy(rr.written).cuda()
            assert torch.equal(w, d_w[tot[2]: tot[2] + w.numel()]), f"kept bytes differ in chunk {k}"
            tot[2] += w.numel()
            if d_o is not None:
                o = torch.from_numpy(rr.other).cuda()
                assert torch.equal(o, d_o[tot[3]: tot[3] + o.numel()]), f"removed bytes differ in chunk {k}"
                tot[3] += o.numel()
            tot[0] += rr.reads_in
            tot[1] += rr.reads_out
    return tuple(tot)


def test_c1_full_size_oracle_parity(ctx):
    """BASELINE configs[0] at full size: 1 M 2x150 pairs + synthetic PAF, --min-len 50 --min-cov 0.5 --min-mapq 50;
    id set, kept / removed bytes and counters equal the oracle's"""
    n = 1_000_000
    paf = synth.gen_paf(n, device="cuda")
    gs = api.IdSet.from_paf(ctx, paf, 50, 0.5, 50)
    oset = orc.set_from_paf(paf.cpu().numpy(), 50, 0.5, 50)
    assert len(gs) == len(oset)
    assert gs.sorted_ids() == oset.sorted_ids()
    for mate in (1, 2):
        _parity_full(ctx, gs, oset, n, mate, False, True, per=250_000)


@pytest.mark.parametrize("reverse", [False, True])
def test_c2_full_size_oracle_parity(ctx, reverse):
    """BASELINE configs[1] at full size: 10 M pairs + Kraken2 reads/report, -T Chordata -D 9606, deplete and -e"""
    from scrubby_b200 import hostlib

    n = 10_000_000
    rep = synth.gen_kraken_report(5000)
    taxids = hostlib.get_taxids_from_report(rep, ["Chordata"], ["9606"])
    otax = orc.taxids_from_report(rep, ["Chordata"], ["9606"])
    assert sorted(t.encode() if isinstance(t, str) else t for t in taxids) == otax.sorted_ids()
    kr = synth.gen_kraken_reads(n, device="cuda")
    gs = api.IdSet.from_reads(ctx, kr, 0, taxids)
    oset = orc.set_from_reads(kr.cpu().numpy(), 0, otax)
    del kr
    assert len(gs) == len(oset)
    # (10 M ids: compare the sets through membership of both complete key lists, sorted on the host)
    assert gs.sorted_ids() == oset.sorted_ids()
    for mate in (1, 2):
        r = _parity_full(ctx, gs, oset, n, mate, reverse, mate == 1)
        assert r.reads_in == n


def test_c3_ont_shape_oracle_parity(ctx):
    """BASELINE configs[2] shape: ONT long reads (lognormal, N50 ~ 10 kb, UUID ids) + a map-ont-style PAF with ~8
    alignments per read grouped by qname (the segmented-reduce stress of alignment.rs:100-108): 200 k reads / 1.6 M PAF
    lines (~3 GB of FASTQ), id set and filtered bytes equal the oracle's"""
    n = 200_000
    fq, lens, uu = synth.gen_ont_fastq(n, device="cuda")
    paf = synth.gen_ont_paf(lens, uu, device="cuda")
    gs = api.IdSet.from_paf(ctx, paf, 50, 0.5, 50)
    oset = orc.set_from_paf(paf.cpu().numpy(), 50, 0.5, 50)
    assert 0.2 * n < len(oset) < 0.8 * n
    assert gs.sorted_ids() == oset.sorted_ids()
    n_in = fq.numel()
    pad = torch.zeros(n_in + 16, dtype=torch.uint8, device="cuda")
    pad[:n_in] = fq
    del fq
    d_w = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    d_o = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    r = api.clean_fastq_dev(ctx, gs, pad, n_in, d_w, d_o)
    assert r.path == 1 and r.reads_in == n
    rec = 1 + 36 + 1 + lens + 1 + 2 + lens + 1
    off = torch.zeros(n + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(rec, 0)
    bounds = [int(off[i]) for i in range(0, n, 20_000)] + [int(off[n])]
    rin, rout, nw, no = _chunks(pad, bounds, oset, False, d_w, d_o)
    assert (rin, rout, nw, no) == (r.reads_in, r.reads_out, r.n_written, r.n_other)


def test_c4_c5_full_size_oracle_parity(ctx):
    """BASELINE configs[3] and [4] at FULL size on one GPU: 100 M 2x150 pairs against the 50 M-id list, then `diff` of
    the input against the depleted output.  The oracle streams over the same generator output in 2.5 M-record chunks on
    the host threads; every chunk's kept bytes must equal the GPU's.  (Skipped when the GPU or the host is too small.)"""
    n = int(os.environ.get("SGPU_TEST_C4_PAIRS", "100000000"))
    free, _ = torch.cuda.mem_get_info()
    need = 1.65 * synth.fastq_size(n) + 6e9
    if free < need:
        pytest.skip(f"needs {need / 1e9:.0f} GB of HBM")
    txt = torch.cat([synth.gen_txt_ids(min(25_000_000, n - s), device="cuda", start=s) for s in range(0, n, 25_000_000)])
    gs = api.IdSet.from_txt(ctx, txt)
    oset = orc.set_from_txt(txt.cpu().numpy())
    assert len(gs) == len(oset)
    del txt
    size = synth.fastq_size(n)
    d_in = torch.empty(size + 64, dtype=torch.uint8, device="cuda")
    d_w = torch.empty(int(size * 0.55), dtype=torch.uint8, device="cuda")
    bounds = _record_bounds(n, 0, 2_500_000)
    for mate in (1, 2):
        synth.gen_fastq(n, mate, device="cuda", out=d_in)
        d_in[size:] = 0
        r = api.clean_fastq_dev(ctx, gs, d_in, size, d_w, None)
        assert r.path == 1 and r.reads_in == n
        rin, rout, nw, _ = _chunks(d_in, bounds, oset, False, d_w, None)
        assert (rin, rout, nw) == (r.reads_in, r.reads_out, r.n_written)
        # config 5: ReadDifference over (input, depleted output): counts against the oracle's counters of this file,
        # the absent ids against the depletion set (every removed read is in it, and only those)
        d = api.diff(ctx, [(d_in[:size], d_w[: r.n_written])])
        assert d[:3] == (n, rout, n - rout)
        assert len(d[3]) == len(oset)
        if mate == 1:
            keys = d[3].keys_dev()
            again = api.IdSet.from_txt(ctx, keys)  # the absent ids == the depletion set: union with it adds nothing
            assert len(again) == len(gs)
            probe = api.diff(ctx, [(d_in[: bounds[1]], d_w[:0])])  # 2.5 M reads vs an empty output: all absent
            assert probe[2] == 2_500_000
            probe[3].free()
            again.free()
            del keys
        d[3].free()


@pytest.mark.parametrize("reverse", [False, True])
def test_long_ids_oracle_parity(ctx, reverse):
    """40-byte Illumina-style read names (fingerprint slot + key arena, the vectorised long-id probe of the fused kernel and
    the word-wise bulk build): set, kept / removed bytes and counters equal the oracle's; ids of 16, 17, 63, 64 and 65
    bytes sit on the edges of that fast path"""
    n = 300_000
    fq = synth.gen_fastq_illumina(n, 1, device="cuda")
    txt = synth.gen_txt_ids_illumina(n, device="cuda")
    gs = api.IdSet.from_txt(ctx, txt)
    oset = orc.set_from_txt(txt.cpu().numpy())
    assert gs.sorted_ids() == oset.sorted_ids()
    n_in = int(fq.numel())
    pad = torch.zeros(n_in + 16, dtype=torch.uint8, device="cuda")
    pad[:n_in] = fq
    d_w = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    d_o = torch.empty(n_in + 64, dtype=torch.uint8, device="cuda")
    r = api.clean_fastq_dev(ctx, gs, pad, n_in, d_w, d_o, reverse)
    o = orc.clean_fastq(fq.cpu().numpy(), oset, reverse, want_bytes=False)
    assert r.path == 1 and (r.reads_in, r.reads_out) == (o.reads_in, o.reads_out)
    assert torch.equal(d_w[: r.n_written].cpu(), torch.from_numpy(o.written))
    assert torch.equal(d_o[: r.n_other].cpu(), torch.from_numpy(o.other))
    # edge lengths, hits and misses, ids that differ only in their last byte
    rng = random.Random(3)
    recs, members = [], []
    for i in range(4000):
        L = rng.choice([15, 16, 17, 31, 32, 33, 47, 48, 63, 64, 65, 90])
        rid = (b"%06d" % i) + bytes(rng.randrange(48, 123) for _ in range(L - 6))
        twin = rid[:-1] + bytes([rid[-1] ^ 1])
        recs.append(b"@" + rid + b" c\n" + b"ACGT" * 30 + b"\n+\n" + b"I" * 120 + b"\n")
        recs.append(b"@" + twin + b"\tc\n" + b"ACGT" * 30 + b"\n+\n" + b"I" * 120 + b"\n")
        if i % 2:
            members.append(rid)
    buf = b"".join(recs)
    g2 = api.IdSet.from_ids(ctx, members)
    got = api.clean_fastq(ctx, g2, buf, reverse)
    want = orc.clean_fastq(buf, orc.OSet.from_ids(members), reverse)
    assert got.path == 1 and got.written == want.written and got.other == want.other
