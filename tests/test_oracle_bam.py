"""Binary BAM evidence (alignment.rs:117-146 through rust-htslib): hand-derived vectors for both restatements of the
reference (C oracle, Python oracle), their agreement on random / corrupted streams, and the host stage that inflates
BGZF.  CPU only; the GPU parity test (tests/test_gpu_parity.py::test_bam_*) uses the same generators."""
import gzip
import os
import random
import struct

import pytest

import bam_build as bb
from oracle import oracle as orc
from oracle import pyoracle as po

E_UTF8, E_BAM = 8, 23


def both(buf, *thr):
    out = []
    for fn, exc, norm in ((orc.set_from_bam, orc.OracleError, lambda s: s.sorted_ids()), (po.ids_from_bam, po.RefError, sorted)):
        try:
            out.append(("ok", norm(fn(buf, *thr))))
        except exc as e:
            out.append(("err", (e.code, e.index)))
    assert out[0] == out[1], (out, buf[:200])
    return out[0]


def test_bam_hand_derived_vectors():
    recs = [
        bb.record(b"r1", cigar="150M"),                      # alen 150
        bb.record(b"r2", cigar="40M110S"),                   # alen 40, cov 0.267 -> out
        bb.record(b"r3", cigar="10S35M15S"),                 # qlen 60, alen 35, cov 0.583 rescues
        bb.record(b"r4", cigar="150M", mapq=49),             # mapq fails
        bb.record(b"r5", flag=4),                            # unmapped
        bb.record(b"r6", cigar="30M10I10D40M70S"),           # M + I = 80; D does not count
        bb.record(b"r7", cigar="40=40X70S"),                 # '=' and 'X' are not Cigar::Match: alen 0 -> out
        bb.record(b"r8", cigar="40M40S", mapq=50),           # cov exactly 0.5, mapq exactly 50: comparisons are >=
        bb.record(b"r9", cigar="", l_seq=0),                 # no CIGAR, no SEQ: alen 0, cov forced to 0.0
        bb.record(b"r5", flag=0x904, cigar="150M"),          # supplementary but flagged unmapped: skipped
        bb.record(b"ra", flag=0x10 | 0x100, cigar="100M50H"),  # secondary reverse: qlen 100 (H is not in SEQ)
    ]
    s = bb.stream(recs)
    assert both(s, 50, 0.5, 50) == ("ok", [b"r1", b"r3", b"r6", b"r8", b"ra"])
    assert both(s) == ("ok", [b"r1", b"r2", b"r3", b"r4", b"r6", b"r7", b"r8", b"r9", b"ra"])  # 0/0/0: every mapped record
    assert both(s, 0, 2.0, 0) == both(s)                       # alen >= 0 always holds
    assert both(s, 1, 2.0, 0) == ("ok", [b"r1", b"r2", b"r3", b"r4", b"r6", b"r8", b"ra"])
    assert both(bb.stream([])) == ("ok", [])
    assert both(bb.stream([], refs=())) == ("ok", [])


def test_bam_qname_rules():
    bad = b"r\xff"
    assert both(bb.stream([bb.record(b"ok", cigar="10M"), bb.record(bad, cigar="10M")])) == ("err", (E_UTF8, 1))
    assert both(bb.stream([bb.record(bad, flag=4), bb.record(b"ok", cigar="10M")])) == ("ok", [b"ok"])  # unmapped: never looked at
    assert both(bb.stream([bb.record("réあ".encode(), cigar="10M")])) == ("ok", ["réあ".encode()])
    assert both(bb.stream([bb.record(b"nonul", cigar="10M", nul=False)])) == ("ok", [b"nonul"])  # htslib appends the NUL
    assert both(bb.stream([bb.record(b"a\x00b", cigar="10M")])) == ("ok", [b"a\x00b"])  # l_read_name - 1 bytes, NULs and all
    assert both(bb.stream([bb.record(b"", cigar="10M")])) == ("ok", [b""])             # read_name "\0": the empty id
    assert both(bb.stream([bb.record(b"", cigar="10M", nul=False)])) == ("err", (E_BAM, 0))  # l_read_name 0


def test_bam_long_cigar_in_cg_tag():
    real = bb.cigar_ops("100M20I30M50S")  # alen 150 of 200
    cg = b"CGBI" + struct.pack("<I", len(real)) + b"".join(struct.pack("<I", v) for v in real)
    other = b"NMi" + struct.pack("<i", 3) + b"MDZ" + b"10A5\x00" + b"XSBc" + struct.pack("<I", 3) + b"\x01\x02\x03"
    fake = [(200 << 4) | 4, (1000 << 4) | 3]  # 200S1000N placeholder
    s = lambda aux, **kw: bb.stream([bb.record(b"long", ops=fake, l_seq=200, aux=aux, **kw)])
    assert both(s(other + cg), 150, 2.0, 0) == ("ok", [b"long"])       # the tag's array is the CIGAR
    assert both(s(other), 1, 2.0, 0) == ("ok", [])                     # no tag: 200S1000N has no M / I
    assert both(s(other + cg, ref_id=-1), 1, 2.0, 0) == ("ok", [])     # bam_tag2cigar needs tid >= 0 and pos >= 0
    assert both(s(other + cg, pos=-1), 1, 2.0, 0) == ("ok", [])
    short = b"CGBI" + struct.pack("<I", 1) + struct.pack("<I", (200 << 4) | 0)
    assert both(s(short), 1, 2.0, 0) == ("ok", [])                     # fewer entries than n_cigar_op: ignored
    assert both(s(b"CGZ" + b"100M\x00" + cg), 1, 2.0, 0) == ("ok", [])  # a CG tag of another type first: not a CIGAR
    assert both(s(b"XXq\x00" + cg), 1, 2.0, 0) == ("ok", [])           # unknown aux type: the walk gives up
    wrap = [(0xFFFFFFF << 4) | 0] * 17  # 17 x (2^28 - 1) M wraps a u32
    assert both(bb.stream([bb.record(b"w", ops=wrap, l_seq=10)]), (17 * 0xFFFFFFF) & 0xFFFFFFFF, 1e30, 0) == ("ok", [b"w"])
    assert both(bb.stream([bb.record(b"w", ops=wrap, l_seq=10)]), ((17 * 0xFFFFFFF) & 0xFFFFFFFF) + 1, 1e30, 0) == ("ok", [])


def test_bam_structural_errors():
    r = [bb.record(b"r%d" % i, cigar="100M") for i in range(5)]
    s = bb.stream(r)
    assert both(s[:-1]) == ("err", (E_BAM, 4))                 # truncated last record
    assert both(s[: len(s) - len(r[4]) + 2]) == ("err", (E_BAM, 4))  # partial block_size
    assert both(s[: len(s) - len(r[4])]) == ("ok", [b"r0", b"r1", b"r2", b"r3"])  # ends on a record boundary
    assert both(b"BAM\x02" + s[4:]) == ("err", (E_BAM, 0))
    assert both(b"") == ("err", (E_BAM, 0)) and both(s[:10]) == ("err", (E_BAM, 0))
    assert both(bb.stream(r[:2] + [struct.pack("<I", 8) + b"\x00" * 8] + r[2:])) == ("err", (E_BAM, 2))  # block_size < 32
    assert both(bb.stream(r[:3] + [bb.record(b"x", cigar="100M", block_size_delta=-60)])) == ("err", (E_BAM, 3))  # fields overrun
    neg = bytearray(bb.record(b"x", cigar="10M"))
    neg[4 + 16: 4 + 20] = struct.pack("<i", -1)               # l_seq < 0
    assert both(bb.stream(r[:1] + [bytes(neg)])) == ("err", (E_BAM, 1))
    # an earlier record's UTF-8 error beats a later truncation
    assert both(bb.stream([bb.record(b"\xff", cigar="10M")] + r)[:-3]) == ("err", (E_UTF8, 0))


def rand_stream(rng, n):
    recs = []
    for i in range(n):
        q = rng.choice([b"r%d" % rng.randrange(40), b"x" * 15, b"x" * 16, b"y" * 40, "qé".encode(), b"r\xff"][: 6 if rng.random() < 0.02 else 5])
        cig = "".join("%d%s" % (rng.choice([1, 5, 40, 50, 100, 151]), rng.choice("MIDNSHP=X")) for _ in range(rng.randrange(0, 6)))
        aux = rng.choice([b"", b"NMi\x01\x00\x00\x00", b"MDZ12A\x00ASC\x07", b"XXBs\x02\x00\x00\x00\x01\x00\x02\x00"])
        recs.append(bb.record(q, flag=rng.choice([0, 0, 0, 16, 4, 256, 2048, 77]), mapq=rng.choice([0, 10, 49, 50, 60, 255]),
                              cigar=cig, aux=aux, ref_id=rng.choice([0, 0, 1, -1]), pos=rng.choice([0, 100, -1]),
                              nul=rng.random() > 0.02))
    return bb.stream(recs, refs=((b"chr1", 1000), (b"chrUn_x", 5)))


@pytest.mark.parametrize("seed", range(25))
def test_bam_two_restatements_agree(seed):
    rng = random.Random(8000 + seed)
    s = rand_stream(rng, rng.choice([1, 20, 200]))
    for thr in [(0, 0.0, 0), (50, 0.5, 50), (100, 2.0, 0), (1 << 40, 0.75, 10)]:
        both(s, *thr)
    for _ in range(10):
        both(s[: rng.randrange(0, len(s))], 50, 0.5, 50)
    for _ in range(20):
        b = bytearray(s)
        for _ in range(rng.choice([1, 1, 3])):
            b[rng.randrange(len(b))] = rng.choice([0, 1, 4, 0x44, 0x7F, 0x80, 0xFF])
        both(bytes(b), 50, 0.5, 50)


def test_host_reader_inflates_bgzf_and_gzip(tmp_path):
    """the host stage (niffler's role): BGZF members in parallel, plain / multi-member gzip serially, raw bytes as is"""
    from scrubby_b200 import hostlib

    rng = random.Random(5)
    s = rand_stream(rng, 3000)
    cases = {"a.bam": bb.bgzf(s), "b.bam": bb.bgzf(s, block=997), "c.gz": gzip.compress(s),
             "d.gz": gzip.compress(s[:1000]) + gzip.compress(s[1000:]), "e.bam": s,
             "f.gz": bb.bgzf(s[:5000]) + gzip.compress(s[5000:])}  # BGZF blocks followed by a plain member: serial path
    for name, data in cases.items():
        p = os.path.join(tmp_path, name)
        with open(p, "wb") as f:
            f.write(data)
        assert hostlib.read_file(p) == s, name
    bad = bytearray(bb.bgzf(s))
    bad[18 + 40] ^= 0x55  # inside the first block's deflate data
    p = os.path.join(tmp_path, "bad.bam")
    with open(p, "wb") as f:
        f.write(bytes(bad))
    with pytest.raises(hostlib.HostError):
        hostlib.read_file(p)
