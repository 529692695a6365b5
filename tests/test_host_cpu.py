"""Host-side pieces of the C++ host that need no GPU: the serde_json / ryu float formatting used for `min_cov` in the
report JSON (report.rs:60 `to_string_pretty`), checked against an independent statement of ryu's "pretty" layout on
top of Python's shortest round-trip digits."""
import math
import random
import struct
from decimal import Decimal

import pytest

from scrubby_b200 import hostlib


def ryu_pretty(v: float) -> str:
    """ryu::Buffer::format_finite (what serde_json writes): shortest round-trip digits d1..dn with decimal exponent,
    laid out as integer.0 / decimal / 0.00ddd for -5 < kk <= 16, else d.ddde±x"""
    if v == 0:
        return "-0.0" if math.copysign(1, v) < 0 else "0.0"
    sign = "-" if v < 0 else ""
    t = Decimal(repr(abs(v))).as_tuple()  # repr: the shortest digits that round-trip (like ryu)
    digits = "".join(map(str, t.digits)).rstrip("0") or "0"
    kk = len(t.digits) + t.exponent  # position of the decimal point relative to the first digit
    n = len(digits)
    if n <= kk <= 16:
        return sign + digits + "0" * (kk - n) + ".0"
    if 0 < kk <= 16:
        return sign + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:
        return sign + "0." + "0" * (-kk) + digits
    if n == 1:
        return sign + digits + "e" + str(kk - 1)
    return sign + digits[0] + "." + digits[1:] + "e" + str(kk - 1)


def test_format_f64_known_values():
    known = {0.5: "0.5", 0.0: "0.0", 1.0: "1.0", 0.1: "0.1", 0.75: "0.75", 100.0: "100.0", 1e16: "1e16", 1e15: "1000000000000000.0",
             1e-5: "0.00001", 1e-6: "1e-6", 1.5e-7: "1.5e-7", 123456.789: "123456.789", 1.2345678901234568e17: "1.2345678901234568e17",
             0.30000000000000004: "0.30000000000000004", 5e-324: "5e-324", 1.7976931348623157e308: "1.7976931348623157e308",
             -2.5: "-2.5"}
    for v, s in known.items():
        assert ryu_pretty(v) == s, v        # the independent statement reproduces ryu's documented outputs
        assert hostlib.format_f64(v) == s, v
    assert hostlib.format_f64(float("nan")) == "null" and hostlib.format_f64(float("inf")) == "null"  # serde_json


@pytest.mark.parametrize("seed", range(4))
def test_format_f64_random_doubles(seed):
    rng = random.Random(seed)
    for _ in range(3000):
        kind = rng.randrange(4)
        if kind == 0:
            v = struct.unpack("<d", struct.pack("<Q", rng.getrandbits(64)))[0]
        elif kind == 1:
            v = round(rng.random(), rng.randrange(1, 6))        # what a user types for --min-cov
        elif kind == 2:
            v = rng.random() * 10 ** rng.randrange(-12, 22)
        else:
            v = float(rng.randrange(0, 10 ** rng.randrange(1, 19)))
        if not math.isfinite(v):
            continue
        assert hostlib.format_f64(v) == ryu_pretty(v), repr(v)


def test_report_string_encoders_match_independent_writers():
    """serde_json escaping against json.dumps(ensure_ascii=False); the csv crate's QuoteStyle::Necessary with a tab
    delimiter against Python's csv writer (QUOTE_MINIMAL)"""
    import csv
    import io
    import json

    rng = random.Random(1)
    alphabet = ["a", "Z", "0", " ", "\t", "\n", "\r", '"', "\\", "/", ",", "\x00", "\x01", "\x1f", "\x7f", "\b", "\f", "é", "あ", "\u2028"]
    cases = ["", "syn.1", 'a"b', "a\tb", "a\nb", "a,b", "tab\there", "\x1f"] + \
        ["".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 12))) for _ in range(2000)]
    for sx in cases:
        raw = sx.encode("utf-8")
        assert hostlib.encode_string(0, raw).decode("utf-8") == json.dumps(sx, ensure_ascii=False), repr(sx)
        out = io.StringIO()
        csv.writer(out, delimiter="\t", quoting=csv.QUOTE_MINIMAL, lineterminator="\n").writerow([sx])
        assert hostlib.encode_string(1, raw).decode("utf-8") + "\n" == out.getvalue(), repr(sx)


def test_cli_help_and_version_need_no_gpu():
    """clap prints help / version and exits 0 before anything is built; usage errors exit 2"""
    import subprocess

    hostlib.build()
    run = lambda *a: subprocess.run([hostlib.CLI, *a], capture_output=True, text=True, timeout=60)
    r = run("--version")
    assert r.returncode == 0 and r.stdout.strip() == "scrubby 1.0.2"
    r = run("--help")
    assert r.returncode == 0 and all(c in r.stdout for c in ("reads", "classifier", "alignment", "diff"))
    for cmd, flag in (("reads", "--index"), ("classifier", "--report"), ("alignment", "--min-mapq"), ("diff", "--read-ids")):
        r = run(cmd, "--help")
        assert r.returncode == 0 and flag in r.stdout and r.stdout.startswith(f"Usage: scrubby {cmd}")
    assert run("reads", "-i", "x", "-A", "-h").returncode == 2       # "-h" is the VALUE of --aligner-args; --index is missing
    assert run("nonsense").returncode == 2 and run().returncode == 2
    assert run("alignment", "-i", "a", "-o", "b").returncode == 2    # --alignment is required


def test_bgzf_member_that_claims_a_huge_payload_is_not_trusted(tmp_path):
    """a crafted BGZF file of tiny members whose ISIZE trailers claim 4 GiB each must not size an allocation:
    the parallel BGZF reader refuses it (the spec caps a block at 64 KiB) and the serial gzip path reports the corrupt
    stream as an error instead of the process dying on bad_alloc"""
    import struct
    import zlib

    def member(payload: bytes, claim: int) -> bytes:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(payload) + co.flush()
        bsize = 12 + 6 + len(body) + 8
        hdr = b"\x1f\x8b\x08\x04" + b"\0" * 4 + b"\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        return hdr + body + struct.pack("<II", zlib.crc32(payload), claim)

    good = member(b"r1\nr2\n", 6) * 40
    p = tmp_path / "ok.gz"
    p.write_bytes(good)
    assert hostlib.read_file(str(p)) == b"r1\nr2\n" * 40
    bad = member(b"r1\n", 0xFFFFFFF0) * 64
    q = tmp_path / "bad.gz"
    q.write_bytes(bad)
    with pytest.raises(Exception):
        hostlib.read_file(str(q))
