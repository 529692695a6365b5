"""FastqCleaner::clean_reads on a plain-gzip FASTQ runs as a pipeline (inflate thread -> shard calls of
the C ABI chunk by chunk -> deflate / write behind it; scrubby_host.cpp: clean_fastq_gz_stream).  The GPU tests run it
through the CLI; HERE the shard call is a stand-in built on the oracle (TEST ONLY), so that the host logic -- chunking,
halo, running newline count, the first chunk's CRLF decision, halo growth, end-of-file rules, error indices, both
writers -- is checked on CPU against the oracle's whole-file result (cleaner.rs:731-760)."""
import ctypes as C
import gzip
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from scrubby_b200 import _lib, hostlib, synth  # noqa: E402

CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                 C.POINTER(C.c_size_t), C.POINTER(_lib.Counts))


def _shard_stand_in(ids, reverse, log):
    """sgpu_clean_fastq_shard's contract (include/scrubby_gpu.h) on the oracle: the records that START in [0, own_len) --
    one that starts exactly at own_len belongs here unless this is the last shard"""
    from oracle import oracle as orc

    def cb(p_in, n_in, own_len, nlb, is_first, is_last, crlf, p_out, cap, p_nout, p_counts):
        buf = C.string_at(p_in, n_in)
        log.append(dict(n_in=n_in, own_len=own_len, nlb=nlb, is_first=is_first, is_last=is_last, crlf=crlf, rc=0))
        p_nout[0] = 0
        pos = 0
        if not is_first:
            for _ in range((3 - nlb % 4) % 4 + 1):
                q = buf.find(b"\n", pos)
                if q < 0:
                    return 0
                pos = q + 1
        start = end = pos
        while end < own_len or (end == own_len and not is_last and end < len(buf)):
            e = end
            for _ in range(4):
                q = buf.find(b"\n", e)
                if q < 0:
                    if is_last:
                        e = len(buf)
                        break
                    log[-1]["rc"] = _lib.SGPU_ERR_HALO
                    return _lib.SGPU_ERR_HALO
                e = q + 1
            end = e
        if start >= end:
            return 0
        r = orc.clean_fastq(buf[start:end], ids, reverse, raise_on_error=False)
        if len(r.written) > cap:
            log[-1]["rc"] = _lib.SGPU_ERR_CAPACITY
            return _lib.SGPU_ERR_CAPACITY
        C.memmove(p_out, r.written, len(r.written))
        p_nout[0] = len(r.written)
        c = p_counts[0]
        c.reads_in, c.reads_out, c.crlf, c.error_record = r.reads_in, r.reads_out, int(r.crlf), r.error_record
        log[-1]["rc"] = r.error
        return r.error

    return CB(cb)


def _stream(path_in, path_out, ids, chunk, halo, reverse=False):
    H = hostlib.load()
    H.scrubby_host_stream_clean.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_size_t, CB, C.POINTER(C.c_int),
                                            C.POINTER(C.c_uint64)]
    log = []
    cb = _shard_stand_in(ids, reverse, log)
    handled, err = C.c_int(-1), C.c_uint64(0)
    rc = H.scrubby_host_stream_clean(str(path_in).encode(), str(path_out).encode(), chunk, halo, cb, C.byref(handled),
                                     C.byref(err))
    return rc, handled.value, err.value, log


def _read_out(path):
    raw = open(path, "rb").read()
    return gzip.decompress(raw) if str(path).endswith(".gz") else raw


def _ids(n):
    from oracle import oracle as orc

    return orc.set_from_txt(synth.gen_txt_ids(n).numpy().tobytes())


@pytest.mark.parametrize("chunk,halo", [(1000, 400), (4096, 512), (50_000, 1 << 16), (1 << 22, 1 << 20)])
@pytest.mark.parametrize("suffix", [".fastq", ".fastq.gz"])
@pytest.mark.parametrize("reverse", [False, True])
def test_gz_stream_matches_the_whole_file(tmp_path, chunk, halo, suffix, reverse):
    from oracle import oracle as orc

    n = 400
    fq = synth.gen_fastq(n, 1).numpy().tobytes()
    ids = _ids(n)
    want = orc.clean_fastq(fq, ids, reverse)
    src = tmp_path / "in.fastq.gz"
    src.write_bytes(gzip.compress(fq, 1))
    dst = tmp_path / ("out" + suffix)
    rc, handled, _, log = _stream(src, dst, ids, chunk, halo, reverse)
    assert (rc, handled) == (0, 1)
    assert _read_out(dst) == want.written
    # the protocol the shard calls saw: first / last flags, a running newline count, the first chunk's line ending
    good = [c for c in log if c["rc"] == 0]
    assert good[0]["is_first"] == 1 and good[0]["crlf"] == -1 and good[-1]["is_last"] == 1
    assert all(c["is_first"] == 0 and c["crlf"] == 0 for c in good[1:]) and all(c["is_last"] == 0 for c in good[:-1])
    assert good[-1]["own_len"] == good[-1]["n_in"]  # the buffer that reaches EOF owns the rest
    pos = 0
    for c in good:
        assert c["nlb"] == fq[:pos].count(b"\n")
        pos += c["own_len"]
    assert pos == len(fq)
    if chunk + halo < len(fq):
        assert len(good) > 1 and all(c["own_len"] == chunk and c["n_in"] >= chunk + halo for c in good[:-1])


@pytest.mark.parametrize("tail", ["multi_member", "no_final_newline", "blank_tail", "crlf", "one_record", "long_records"])
def test_gz_stream_edge_inputs(tmp_path, tail):
    from oracle import oracle as orc

    n = 120
    fq = synth.gen_fastq(n, 2).numpy().tobytes()
    ids = _ids(n)
    if tail == "no_final_newline":
        fq = fq[:-1]
    elif tail == "blank_tail":
        fq += b"\n\n"
    elif tail == "crlf":
        fq = fq.replace(b"\n", b"\r\n")
    elif tail == "one_record":
        fq = fq[: fq.index(b"\n@syn.", 1) + 1]
    elif tail == "long_records":  # longer than the first halo (and than a whole chunk): the halo grows until one fits
        recs = [b"@syn.%d/2\n" % i + b"ACGT" * k + b"\n+\n" + b"IIII" * k + b"\n" for i, k in enumerate((3, 900, 5, 2500, 7, 40, 1))]
        fq = b"".join(recs)
    gz = gzip.compress(fq, 6)
    if tail == "multi_member":  # members cut in the middle of a record
        gz = gzip.compress(fq[:7001]) + gzip.compress(fq[7001:20000]) + gzip.compress(fq[20000:])
    want = orc.clean_fastq(fq, ids)
    assert want.reads_in and want.crlf == (tail == "crlf")
    src = tmp_path / "in.fq.gz"
    src.write_bytes(gz)
    for chunk, halo in ((700, 350), (3000, 64), (1 << 20, 1 << 16)):
        dst = tmp_path / f"out_{chunk}.fq"
        rc, handled, _, log = _stream(src, dst, ids, chunk, halo)
        assert (rc, handled) == (0, 1), (tail, chunk)
        assert _read_out(dst) == want.written, (tail, chunk)
        if tail == "crlf" and chunk < len(fq):
            assert [c["crlf"] for c in log if c["rc"] == 0][:2] == [-1, 1]
        if tail == "long_records" and chunk == 3000:
            assert any(c["rc"] == _lib.SGPU_ERR_HALO for c in log)


def test_gz_stream_parse_error_in_a_later_chunk(tmp_path):
    """the records before the failing one are written, the error carries the record's index in the FILE"""
    from oracle import oracle as orc

    n = 200
    fq = bytearray(synth.gen_fastq(n, 1).numpy().tobytes())
    starts = [0]
    for _ in range(4 * 150):
        starts.append(fq.index(b"\n", starts[-1]) + 1)
    plus = starts[4 * 150 - 2]  # the separator line of record 149
    assert fq[plus: plus + 1] == b"+"
    fq[plus] = ord("-")
    fq = bytes(fq)
    ids = _ids(n)
    want = orc.clean_fastq(fq, ids, raise_on_error=False)
    assert want.error == 4 and want.error_record == 149
    src = tmp_path / "bad.fastq.gz"
    src.write_bytes(gzip.compress(fq))
    for chunk in (2000, 1 << 20):
        dst = tmp_path / f"out_{chunk}.fastq.gz"
        rc, handled, err, _ = _stream(src, dst, ids, chunk, 700)
        assert rc >= 100 and err == 149
        assert _read_out(dst) == want.written


def test_gz_stream_leaves_other_inputs_to_the_whole_file_path(tmp_path):
    fq = synth.gen_fastq(50, 1).numpy().tobytes()
    ids = _ids(50)
    cases = {
        "plain.fastq": fq,                                          # not gzip
        "fasta.fa.gz": gzip.compress(b">a\nACGT\n>b\nAC\nGT\n"),    # FASTA: no shard entry point
        "empty.fastq.gz": gzip.compress(b""),                       # nothing inside: the empty-input rule
        "tiny.gz": b"\x1f\x8b\x08",                                 # niffler: FileTooShort
        "junk.fastq.gz": gzip.compress(b"hello\n"),                 # neither '@' nor '>': UnknownFormat, no output file
    }
    for name, raw in cases.items():
        src = tmp_path / name
        src.write_bytes(raw)
        dst = tmp_path / (name + ".out")
        rc, handled, _, log = _stream(src, dst, ids, 1000, 100)
        assert (rc, handled) == (0, 0), name
        assert not log and not dst.exists(), name


def test_gz_stream_corrupt_member(tmp_path):
    fq = synth.gen_fastq(300, 1).numpy().tobytes()
    gz = bytearray(gzip.compress(fq))
    gz[len(gz) // 2] ^= 0xFF
    gz[len(gz) // 2 + 1] ^= 0xFF
    src = tmp_path / "corrupt.fastq.gz"
    src.write_bytes(bytes(gz))
    rc, handled, _, _ = _stream(src, tmp_path / "o.fastq", _ids(300), 2000, 500)
    # (a stream reader meets whichever comes first: the inflate error, or a parse error in the bytes decoded before it
    #  -- as the reference's needletail-over-flate2 reader does)
    assert rc >= 100
    # garbage after a complete member is an error too (as in read_file)
    src.write_bytes(gzip.compress(fq) + b"garbage-after-the-member")
    rc, handled, _, _ = _stream(src, tmp_path / "o2.fastq", _ids(300), 2000, 500)
    assert rc == 100 + hostlib.KIND_IO


@pytest.mark.parametrize("kind", ["bgzf", "bgzf_small_blocks", "bgzf_then_gzip", "gzip_then_bgzf", "bgzf_truncated", "bgzf_corrupt"])
def test_bgzf_stream(tmp_path, monkeypatch, kind):
    """BGZF input (bgzip): members cut out by BSIZE and inflated batch by batch on all threads; a member that is not BGZF
    (appended plain gzip, a truncated tail) hands the rest of the file to the serial reader -- the same bytes as the
    whole-file reader (scrubby::read_file) gives, through the same chunked shard calls"""
    from bam_build import bgzf
    from oracle import oracle as orc

    n = 2500
    fq = synth.gen_fastq(n, 1).numpy().tobytes()
    ids = _ids(n)
    cut = len(fq) // 2 + 77
    raw = {"bgzf": bgzf(fq), "bgzf_small_blocks": bgzf(fq, 700),
           "bgzf_then_gzip": bgzf(fq[:cut])[:-28] + gzip.compress(fq[cut:]),     # (without the EOF marker in between)
           "gzip_then_bgzf": gzip.compress(fq[:cut]) + bgzf(fq[cut:]),
           "bgzf_truncated": bgzf(fq)[: len(bgzf(fq)) * 2 // 3],
           "bgzf_corrupt": bgzf(fq)}[kind]
    if kind == "bgzf_corrupt":
        b = bytearray(raw)
        b[len(b) // 2] ^= 0x55
        raw = bytes(b)
    src = tmp_path / "in.fastq.gz"
    src.write_bytes(raw)
    try:
        whole = hostlib.read_file(str(src))   # what the whole-file path would filter
    except hostlib.HostError:
        whole = None
    assert (whole is None) == (kind == "bgzf_corrupt")
    if kind in ("bgzf", "bgzf_small_blocks", "bgzf_then_gzip", "gzip_then_bgzf"):
        assert whole == fq
    for piece, chunk, halo in ((1 << 14, 30_000, 2000), (1 << 23, 1 << 18, 1 << 14), (1 << 12, 5000, 800)):
        monkeypatch.setenv("SCRUBBY_BGZF_PIECE", str(piece))
        dst = tmp_path / f"out_{piece}.fastq"
        rc, handled, err, log = _stream(src, dst, ids, chunk, halo)
        assert handled == 1 or rc
        if whole is None:
            assert rc == 100 + hostlib.KIND_IO or rc >= 100
            continue
        want = orc.clean_fastq(whole, ids, raise_on_error=False)
        if want.error:
            assert rc >= 100 and err == want.error_record
        else:
            assert rc == 0
        assert _read_out(dst) == want.written
        assert len([c for c in log if c["rc"] == 0]) > 1
    monkeypatch.setenv("SCRUBBY_NO_BGZF_STREAM", "1")
    rc, handled, _, log = _stream(src, tmp_path / "no.fastq", ids, 30_000, 2000)
    if kind != "gzip_then_bgzf":  # (that one starts with a plain member: it is the plain-gzip stream's case)
        assert (rc, handled) == (0, 0) and not log
