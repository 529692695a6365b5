"""Randomised differential tests of the two INDEPENDENT restatements of the reference (oracle/scrubby_oracle.c and
oracle/pyoracle.py) on adversarial inputs: the reference ships no tests or fixtures, so the oracle is pinned
by the hand-derived golden vectors and by these two restatements agreeing on everything random generators can reach --
results, error classes and error indices alike.  CPU only."""
import random

import pytest

from oracle import oracle as orc
from oracle import pyoracle as po

def _err_of(e: "orc.OracleError"):
    return (getattr(e, "code", None) or getattr(e, "status", None), getattr(e, "index", None))


def _both(cfn, pfn, cargs, pargs, norm_c, norm_p=lambda x: x):
    """run both restatements; equal results, or equal (error class, error index)"""
    try:
        rc = ("ok", norm_c(cfn(*cargs)))
    except orc.OracleError as e:
        rc = ("err", _err_of(e))
    try:
        rp = ("ok", norm_p(pfn(*pargs)))
    except po.RefError as e:
        rp = ("err", (e.code, e.index))
    assert rc == rp, (rc if rc[0] == "err" else "ok", rp if rp[0] == "err" else "ok", cargs[0][:300])
    return rc


def test_oracle_error_exposes_code_and_index():
    with pytest.raises(orc.OracleError) as ei:
        orc.set_from_paf(b"r0\t1\t0\t1\t+\tt\t1\t0\t1\t1\t1\t0\na\n")  # fields[1] of line 1 is out of bounds
    assert _err_of(ei.value) == (po.E_PANIC, 1)
    with pytest.raises(orc.OracleError) as ei:
        orc.set_from_paf(b"a\tb\n")  # struct fields are evaluated in order: the parse error comes before the panic
    assert _err_of(ei.value) == (po.E_PAFINT, 0)


# ---------------------------------------------------------------------------------------------- FASTQ
def _rand_fastq(rng, n_rec, pool):
    out = bytearray()
    crlf_first = rng.random() < 0.3
    for i in range(n_rec):
        style = rng.randrange(10)
        rid = rng.choice(pool)
        # White_Space per Rust (split points): VT, NBSP, EM SPACE, NEL, IDEOGRAPHIC SPACE; NOT White_Space: FS, ZWSP, BOM
        desc = rng.choice(["", " d", "\t1:N:0", " a b c", "\x0bvt", " \u00fc", " \r x", "\u2003em", "\u00a0nb", "\u0085nel",
                           "\u3000id", "\x1cfs", "\u200bzw", "\ufeffbom"]).encode()
        lead = rng.choice([b"", b"", b"", b" ", "\u00a0".encode(), "\u2003".encode()])
        L = rng.choice([0, 1, 2, 15, 16, 17, 31, 64, 150])
        seq = bytes(rng.choice(b"ACGTN") for _ in range(L))
        qual = bytes(rng.randrange(33, 74) for _ in range(L))
        if L and style == 1:
            qual = b"@" + qual[1:]
        if L and style == 2:
            qual = b"+" + qual[1:]
        e = b"\r\n" if (crlf_first and i == 0) or style == 3 else b"\n"
        sep = b"+" + (rid + desc if style == 4 else b"")
        out += b"@" + lead + rid + desc + e + seq + e + sep + e + qual + e
    return bytes(out)


@pytest.mark.parametrize("seed", range(8))
def test_fastq_clean_two_restatements_agree(seed):
    rng = random.Random(1000 + seed)
    pool = [f"id{k}".encode() for k in range(40)] + [b"x" * 15, b"x" * 16, b"y" * 40, "é1".encode(), "réあ".encode()]
    ids = rng.sample(pool, 18)
    oset, pset = orc.OSet.from_ids(ids), set(ids)
    base = _rand_fastq(rng, 120, pool)
    cases = [base + t for t in (b"", b"\n", b"\n\n", b"\r\n", b"\r\n\n", b" \n")]
    cases.append(base[:-1])
    cases += [base[: rng.randrange(1, len(base))] for _ in range(25)]
    for _ in range(25):
        b2 = bytearray(base)
        for _ in range(rng.choice([1, 1, 2])):
            b2[rng.randrange(len(b2))] = rng.choice(b"\n@+\rA \xff\xc3")
        cases.append(bytes(b2))
    cases += [b"", b"@", b"@a\n", b"@a\nA\n+\nI", b">a\nACGT\n", b"Xa\nA\n+\nI\n", b"\n\n\n\n\n", b"@a\nA\n+\nI\n\n\n@"]
    for buf in cases:
        for reverse in (False, True):
            r = _both(lambda b: orc.clean_fastq(b, oset, reverse), lambda b: po.clean_fastq(b, pset, reverse),
                      (buf,), (buf,), lambda c: (c.written, c.other, c.reads_in, c.reads_out))
            if r[0] == "ok" and r[1][2]:
                # the output of a run is itself canonical for the reference: a fixed point in both restatements
                w = r[1][0]
                again = orc.clean_fastq(w, oset, reverse)
                assert (again.written, again.reads_in, again.reads_out) == (w, r[1][3], r[1][3])


@pytest.mark.parametrize("seed", range(4))
def test_diff_two_restatements_agree(seed):
    rng = random.Random(2000 + seed)
    pool = [f"q{k}".encode() for k in range(60)] + [b"z" * 16, b"z" * 33]
    pairs = []
    for _ in range(2):
        fin = _rand_fastq(rng, 80, pool)
        keep = orc.OSet.from_ids(rng.sample(pool, 25))
        try:
            fout = orc.clean_fastq(fin, keep).written
        except orc.OracleError:
            fout = b""
        pairs.append((fin, fout))
    _both(orc.diff, po.diff, (pairs,), (pairs,), lambda d: (d[0], d[1], d[2], d[3].sorted_ids()),
          lambda d: (d[0], d[1], d[2], sorted(d[3])))


# ---------------------------------------------------------------------------------------------- PAF
def _rand_uint(rng, bits=64):
    r = rng.random()
    if r < 0.70:
        return str(rng.choice([0, 1, 40, 49, 50, 51, 75, 100, 150, 151, 255, 256, 10 ** 6]))
    if r < 0.78:
        return "+" + str(rng.randrange(300))
    if r < 0.84:
        return str((1 << bits) - 1 - rng.randrange(2))
    return rng.choice(["", " 15", "15 ", "-1", "1e3", "0x10", "١٢", str(1 << bits), "+", "-", "1_0", "+-1", "00012"])


def _rand_paf(rng, n):
    lines = []
    for i in range(n):
        q = rng.choice([f"r{rng.randrange(30)}", "r 1", "", "é", " r2", "r3 "])
        f = [q, _rand_uint(rng), _rand_uint(rng), _rand_uint(rng), rng.choice("+-"), "chr1", _rand_uint(rng),
             _rand_uint(rng), _rand_uint(rng), _rand_uint(rng), _rand_uint(rng), _rand_uint(rng, 8)]
        if rng.random() < 0.5:
            f += ["tp:A:P", "cm:i:10"]
        if rng.random() < 0.03:
            f = f[: rng.randrange(0, 12)]
        lines.append("\t".join(f) + rng.choice(["\n", "\n", "\n", "\r\n"]))
    s = "".join(lines)
    if rng.random() < 0.3:
        s = s.rstrip("\n")
    return s.encode()


@pytest.mark.parametrize("seed", range(40))
def test_paf_two_restatements_agree(seed):
    rng = random.Random(3000 + seed)
    # error-free prefixes matter most (an erroring file only pins the first error): mostly clean generators
    n = rng.choice([3, 10, 40])
    buf = _rand_paf(rng, n)
    if seed % 2 == 0:  # make every integer valid: the predicate arithmetic (wrapping alen, f64 coverage) is what differs
        rng2 = random.Random(seed)
        rows = []
        for i in range(200):
            qlen, qs, qe = rng2.choice([0, 1, 60, 80, 150]), rng2.randrange(0, 160), rng2.randrange(0, 160)
            rows.append(f"r{i % 37}\t{qlen}\t{qs}\t{qe}\t+\tt\t1000\t0\t100\t90\t100\t{rng2.choice([0, 10, 49, 50, 60, 255])}\n")
        buf = "".join(rows).encode()
    for ml, mc, mq in [(0, 0.0, 0), (50, 0.5, 50), (0, 0.5, 0), (50, 0.0, 0), (1 << 63, 2.0, 0), (150, 1.0, 255)]:
        _both(orc.set_from_paf, po.ids_from_paf, (buf, ml, mc, mq), (buf, ml, mc, mq), lambda s: s.sorted_ids(), sorted)
    bad = buf[: len(buf) // 2] + b"\xff\xfe" + buf[len(buf) // 2:]
    _both(orc.set_from_paf, po.ids_from_paf, (bad,), (bad,), lambda s: s.sorted_ids(), sorted)


# ---------------------------------------------------------------------------------------------- TXT / Kraken2 / Metabuli
@pytest.mark.parametrize("seed", range(10))
def test_txt_and_reads_two_restatements_agree(seed):
    rng = random.Random(4000 + seed)
    toks = ["r1", "r 1", " r1", "r1 ", "", "é", "x" * 16, "x" * 40, "a\tb", "r\u2003", "\x1cr"]
    txt = "".join(rng.choice(toks) + rng.choice(["\n", "\n", "\r\n"]) for _ in range(60))
    if rng.random() < 0.5:
        txt = txt.rstrip("\n")
    _both(orc.set_from_txt, po.ids_from_txt, (txt.encode(),), (txt.encode(),), lambda s: s.sorted_ids(), sorted)
    taxids = [b"9606", b"7711", b"09606", b"+9606", b"Homo sapiens (taxid 9606)", b"0", "é".encode()]
    c_tax, p_tax = orc.OSet.from_ids(taxids), set(taxids)
    for style, need in ((0, 5), (1, 7)):
        lines = []
        for i in range(80):
            pad = lambda s: rng.choice(["", " ", "\u00a0", "\u2003"]) + s + rng.choice(["", " ", "\u3000", "\u0085"])
            rid = pad(rng.choice(["r%d" % rng.randrange(20), "é", "x" * 20, ""]))
            tid = pad(rng.choice(["9606", "7711", "09606", "+9606", "562", "0", "Homo sapiens (taxid 9606)", "é", "\x1c9606"]))
            f = [rng.choice("CU"), rid, tid] + ["150|150", "0:1 9606:5"] + ["x", "y", "z"][: rng.randrange(0, 4)]
            if rng.random() < 0.02:
                f = f[: rng.randrange(0, need)]
            lines.append("\t".join(f) + rng.choice(["\n", "\n", "\r\n"]))
        buf = "".join(lines).encode()
        _both(lambda b: orc.set_from_reads(b, style, c_tax), lambda b: po.ids_from_reads(b, style, p_tax),
              (buf,), (buf,), lambda s: s.sorted_ids(), sorted)


# ---------------------------------------------------------------------------------------------- report state machine
RANKS_K = ["U", "R", "R1", "D", "D1", "K", "K1", "K2", "P", "P1", "C", "C2", "O", "F", "F1", "G", "S", "S1", "S2", "-", "X"]
RANKS_M = ["no rank", "superkingdom", "clade", "kingdom", "phylum", "subphylum", "class", "order", "family", "genus",
           "species", "subspecies", "strain"]


@pytest.mark.parametrize("seed", range(30))
def test_report_state_machine_two_restatements_agree(seed):
    rng = random.Random(5000 + seed)
    ranks = RANKS_K if seed % 3 else RANKS_M
    names = ["root", "Bacteria", "Eukaryota", "Chordata", "Homo", "Homo sapiens", "Mammalia", "Aves", "n%d" % seed]
    lines = []
    for i in range(rng.choice([5, 30, 120])):
        name = rng.choice(names + ["t%d" % i] * 6)
        tid = rng.choice(["9606", "7711", "2", "1", str(100 + i)])
        direct = rng.choice(["0", "0", "1", "17", "+3"])
        if rng.random() < 0.01:
            direct = rng.choice(["", "x", "-1"])
        f = ["%.2f" % rng.random(), str(rng.randrange(1000)), direct, rng.choice(ranks), tid,
             " " * rng.randrange(0, 6) + name + rng.choice(["", " ", "\u00a0", "\u2003"])]
        if rng.random() < 0.01:
            f = f[: rng.randrange(0, 6)]
        lines.append("\t".join(f) + rng.choice(["\n", "\n", "\r\n"]))
    buf = "".join(lines).encode()
    for taxa, direct in [(["Chordata"], ["9606"]), (["7711"], []), ([], ["root", "2"]), (["Eukaryota", " Homo "], ["Aves"]),
                         (["root"], []), (["n%d" % seed], ["t3"])]:
        _both(orc.taxids_from_report, po.taxids_from_report, (buf, taxa, direct), (buf, taxa, direct),
              lambda s: s.sorted_ids(), sorted)


# ---------------------------------------------------------------------------------------------- product host code
def _rand_report(rng, seed):
    ranks = RANKS_K if seed % 3 else RANKS_M
    names = ["root", "Bacteria", "Eukaryota", "Chordata", "Homo", "Homo sapiens", "Mammalia", "Aves", "n%d" % seed]
    lines = []
    for i in range(rng.choice([5, 30, 120])):
        name = rng.choice(names + ["t%d" % i] * 6)
        tid = rng.choice(["9606", "7711", "2", "1", str(100 + i)])
        direct = rng.choice(["0", "0", "1", "17", "+3"])
        if rng.random() < 0.01:
            direct = rng.choice(["", "x", "-1"])
        f = ["%.2f" % rng.random(), str(rng.randrange(1000)), direct, rng.choice(ranks), tid,
             " " * rng.randrange(0, 6) + name + rng.choice(["", " ", " ", " "])]
        if rng.random() < 0.01:
            f = f[: rng.randrange(0, 6)]
        lines.append("\t".join(f) + rng.choice(["\n", "\n", "\r\n"]))
    return "".join(lines).encode()


@pytest.mark.parametrize("seed", range(30))
def test_host_report_state_machine_matches_oracle(seed):
    """the C++ host's get_taxids_from_report (classifier.rs:124-252; it stays on the host by design) against
    the oracle on random reports: same taxid sets, same error class and line"""
    from scrubby_b200 import hostlib

    kinds = {hostlib.KIND_IO: po.E_IO, hostlib.KIND_KR_PARENT: po.E_KR_PARENT, hostlib.KIND_KR_READS: po.E_KR_READS,
             hostlib.KIND_KR_DIRECT: po.E_KR_DIRECT, hostlib.KIND_WOULD_PANIC: po.E_PANIC}
    rng = random.Random(5000 + seed)
    buf = _rand_report(rng, seed)
    for taxa, direct in [(["Chordata"], ["9606"]), (["7711"], []), ([], ["root", "2"]), (["Eukaryota", " Homo "], ["Aves"]),
                         (["root"], []), (["n%d" % seed], ["t3"])]:
        try:
            want = ("ok", orc.taxids_from_report(buf, taxa, direct).sorted_ids())
        except orc.OracleError as e:
            want = ("err", _err_of(e))
        try:
            got = ("ok", hostlib.get_taxids_from_report(buf, taxa, direct))
        except hostlib.HostError as e:
            got = ("err", (kinds.get(e.kind, -e.kind), e.line))
        assert got == want, (taxa, direct)


# ---------------------------------------------------------------------------------------------- FASTA input
FASTA_CASES = [  # (input, ids in the set, reverse, written, other) -- derived by hand from needletail's fasta reader rules
    (b">a x\nAC\nGT\n>b\nTT\n>c\n>d\nAA", [b"b"], False, b">a x\nAC\nGT\n>c\n\n>d\nAA\n", b">b\nTT\n"),   # multi-line kept verbatim
    (b">a x\nAC\nGT\n>b\nTT\n>c\n>d\nAA", [b"b"], True, b">b\nTT\n", b">a x\nAC\nGT\n>c\n\n>d\nAA\n"),
    (b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n", [b"b"], False, b">a\r\nAC\r\nGT\r\n", b">b\r\nTT\r\n"),               # CRLF file
    (b">a\nAC\n\n", [], False, b">a\nAC\n\n", b""),                                                            # blank last line
    (b">a\n>b\r\nAC\r\n", [], False, b">a\n\n>b\r\nAC\r\n", b""),          # the line ending is decided by the first record that holds a newline
    (b">a\nAC\r\n>b\n\nTT\r", [b"a"], True, b">a\nAC\n", b">b\n\nTT\n"),   # one trailing CR trimmed from the id and from raw_seq
    (b">  a b\nAC\n", [b"a"], False, b"", b">  a b\nAC\n"),                 # get_id: first whitespace-separated token
]


def test_fasta_hand_derived_cases():
    for buf, ids, rev, w, o in FASTA_CASES:
        r = orc.clean_fastq(buf, orc.OSet.from_ids(ids), rev)
        assert (r.written, r.other) == (w, o), buf
        p = po.clean_fastq(buf, set(ids), rev)
        assert (p[0], p[1]) == (w, o), buf
    for buf, err in ((b">abcd", (po.E_END, 0)), (b">a\nAC\n>b\n", (po.E_END, 1)), (b">a\nAC\n>\nTT\n", (po.E_HEADER, 1)),
                     (b">a\nAC\n>\xff\nTT\n", (po.E_UTF8, 1))):
        with pytest.raises(orc.OracleError) as ei:
            orc.clean_fastq(buf, orc.OSet.from_ids([]))
        assert _err_of(ei.value) == err, buf
        with pytest.raises(po.RefError) as pi:
            po.clean_fastq(buf, set())
        assert (pi.value.code, pi.value.index) == err, buf


def rand_fasta(rng, n):
    parts = []
    for i in range(n):
        rid = rng.choice([b"q%d" % rng.randrange(6), b"b", b" x", b"a b", b"x" * 16, b"\xc3\xa9", b"", b"\xff"][: 8 if rng.random() < 0.03 else 6])
        e_ = rng.choice([b"\n", b"\n", b"\r\n"])
        lines = [bytes(rng.choice(b"ACGT>\r ") for _ in range(rng.choice([0, 1, 5, 60]))) for _ in range(rng.randrange(0, 4))]
        parts.append(b">" + rid + e_ + b"".join(l + e_ for l in lines))
    buf = b"".join(parts)
    if rng.random() < 0.4:
        buf = buf.rstrip(b"\n")
    if rng.random() < 0.2:
        buf = buf[: rng.randrange(1, len(buf) + 1)]
    return buf


@pytest.mark.parametrize("seed", range(10))
def test_fasta_two_restatements_agree(seed):
    rng = random.Random(9000 + seed)
    ids = [b"b", b"q3", b"x" * 16]
    oset, pset = orc.OSet.from_ids(ids), set(ids)
    for _ in range(150):
        buf = rand_fasta(rng, rng.randrange(1, 9))
        rev = rng.random() < 0.5
        _both(lambda b: orc.clean_fastq(b, oset, rev), lambda b: po.clean_fastq(b, pset, rev), (buf,), (buf,),
              lambda c: (c.written, c.other, c.reads_in, c.reads_out))
    pairs = []
    for _ in range(2):
        fin = rand_fasta(rng, 30)
        try:
            fout = orc.clean_fastq(fin, oset).written
        except orc.OracleError:
            fout = b""
        pairs.append((fin, fout))
    _both(orc.diff, po.diff, (pairs,), (pairs,), lambda d: (d[0], d[1], d[2], d[3].sorted_ids()),
          lambda d: (d[0], d[1], d[2], sorted(d[3])))
