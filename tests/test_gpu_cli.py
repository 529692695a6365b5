"""End-to-end runs of the `scrubby` CLI host (scrubby_b200/host, same subcommands and flags as the
reference's terminal.rs) on a GPU: output FASTQ bytes, report JSON and read-id TSV against the CPU oracle."""
import gzip
import json
import os
import subprocess

import pytest

from oracle import oracle as orc
from scrubby_b200 import hostlib, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cli():
    hostlib.build()
    assert os.path.exists(hostlib.CLI)
    return hostlib.CLI


def _run(cli, *args, ok=True):
    r = subprocess.run([cli, *map(str, args)], capture_output=True, text=True, timeout=300)
    assert (r.returncode == 0) == ok, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    return r


def _write(p, b):
    with open(p, "wb") as f:
        f.write(b)
    return str(p)


def _tsv_ids(p):
    rows = open(p).read().split("\n")
    assert rows[0] == "id"
    return sorted(x for x in rows[1:] if x)


@pytest.fixture(scope="module")
def data(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    n = 20_000
    fq = [synth.gen_fastq(n, m).numpy().tobytes() for m in (1, 2)]
    paf = synth.gen_paf(n).numpy().tobytes()
    kr = synth.gen_kraken_reads(n).numpy().tobytes()
    rep = synth.gen_kraken_report(2000)
    files = dict(d=d, n=n, fq=fq, paf=paf, kr=kr, rep=bytes(rep),
                 r1=_write(d / "r1.fq", fq[0]), r2=_write(d / "r2.fq", fq[1]), paf_p=_write(d / "aln.paf", paf),
                 kr_p=_write(d / "reads.kraken", kr), rep_p=_write(d / "report.txt", bytes(rep)))
    return files


@pytest.mark.parametrize("extract", [False, True])
def test_cli_alignment_paf(cli, data, extract):
    d = data["d"]
    o1, o2, js, tsv = d / f"o1_{extract}.fq", d / f"o2_{extract}.fq", d / f"rep_{extract}.json", d / f"ids_{extract}.tsv"
    args = ["alignment", "-i", data["r1"], data["r2"], "-o", o1, o2, "-a", data["paf_p"], "--min-len", 50, "--min-cov", 0.5,
            "--min-mapq", 50, "-j", js, "-r", tsv]
    if extract:
        args.append("-e")
    _run(cli, *args)
    oset = orc.set_from_paf(data["paf"], 50, 0.5, 50)
    want = [orc.clean_fastq(f, oset, extract) for f in data["fq"]]
    assert open(o1, "rb").read() == want[0].written and open(o2, "rb").read() == want[1].written
    rep = json.load(open(js))
    od = orc.diff([(data["fq"][i], want[i].written) for i in range(2)])
    assert list(rep.keys()) == ["version", "date", "command", "input", "output", "reads_in", "reads_out", "reads_removed",
                                "reads_extracted", "settings"]
    assert (rep["reads_in"], rep["reads_out"]) == (od[0], od[1])
    assert (rep["reads_removed"], rep["reads_extracted"]) == ((0, od[2]) if extract else (od[2], 0))  # report.rs:44-45
    assert rep["settings"]["min_len"] == 50 and rep["settings"]["extract"] is extract
    assert _tsv_ids(tsv) == sorted(x.decode() for x in od[3].sorted_ids())


def test_cli_classifier_and_diff(cli, data):
    d = data["d"]
    o1, o2 = d / "c1.fq", d / "c2.fq"
    _run(cli, "classifier", "-i", data["r1"], data["r2"], "-o", o1, o2, "--report", data["rep_p"], "--reads", data["kr_p"],
         "-c", "kraken2", "-T", "Chordata", "-D", "9606")
    tax = orc.taxids_from_report(data["rep"], ["Chordata"], ["9606"])
    oset = orc.set_from_reads(data["kr"], 0, tax)
    want = [orc.clean_fastq(f, oset, False) for f in data["fq"]]
    assert open(o1, "rb").read() == want[0].written and open(o2, "rb").read() == want[1].written
    js, tsv = d / "diff.json", d / "diff.tsv"
    _run(cli, "diff", "-i", data["r1"], data["r2"], "-o", o1, o2, "-j", js, "-r", tsv)
    od = orc.diff([(data["fq"][i], want[i].written) for i in range(2)])
    assert json.load(open(js)) == {"reads_in": od[0], "reads_out": od[1], "difference": od[2]}  # utils.rs:180-187
    assert _tsv_ids(tsv) == sorted(x.decode() for x in od[3].sorted_ids())


def test_cli_gzip_in_and_out(cli, data):
    """gz is a host stage (niffler's role): parity is on the DEcompressed bytes"""
    d = data["d"]
    g1 = d / "r1.fq.gz"
    with gzip.open(g1, "wb", compresslevel=1) as f:
        f.write(data["fq"][0])
    o1 = d / "g1.fq.gz"
    _run(cli, "alignment", "-i", g1, "-o", o1, "-a", data["paf_p"], "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50)
    want = orc.clean_fastq(data["fq"][0], orc.set_from_paf(data["paf"], 50, 0.5, 50), False)
    assert gzip.open(o1, "rb").read() == want.written


def test_cli_sam_format_and_errors(cli, data):
    d = data["d"]
    # a SAM file that removes every third read
    hdr = "@HD\tVN:1.6\n@SQ\tSN:chr1\tLN:1000\n"
    lines = [f"syn.{i}\t0\tchr1\t10\t60\t150M\t*\t0\t0\t{'A' * 150}\t{'I' * 150}" for i in range(0, data["n"], 3)]
    sam = (hdr + "\n".join(lines) + "\n").encode()
    sp = _write(d / "aln.sam", sam)
    o1 = d / "s1.fq"
    _run(cli, "alignment", "-i", data["r1"], "-o", o1, "-a", sp, "--min-len", 100)
    want = orc.clean_fastq(data["fq"][0], orc.set_from_sam(sam, 100, 0.0, 0), False)
    assert open(o1, "rb").read() == want.written and want.reads_out < want.reads_in
    # the same file with an explicit format and a name that has no known extension
    sp2 = _write(d / "aln.dat", sam)
    o2 = d / "s2.fq"
    _run(cli, "alignment", "-i", data["r1"], "-o", o2, "-a", sp2, "-f", "sam", "--min-len", 100)
    assert open(o2, "rb").read() == want.written
    # unknown extension without --format: AlignmentInputFormatNotRecognized
    assert _run(cli, "alignment", "-i", data["r1"], "-o", d / "x.fq", "-a", sp2, ok=False).returncode != 0
    # sam / bam / cram share one htslib reader that looks at the CONTENT (alignment.rs:45): SAM text under -f bam works
    o3 = d / "s3.fq"
    _run(cli, "alignment", "-i", data["r1"], "-o", o3, "-a", sp2, "-f", "bam", "--min-len", 100)
    assert open(o3, "rb").read() == want.written
    # mismatched input / output counts (scrubby.rs:760-779)
    assert _run(cli, "alignment", "-i", data["r1"], data["r2"], "-o", d / "x.fq", "-a", data["paf_p"], ok=False).returncode != 0


def test_cli_bam_alignment(cli, data):
    """binary BAM evidence, BGZF-compressed as samtools writes it (host stage: parallel BGZF inflate; GPU: one thread
    per record), by extension and by --format; truncated BAM fails like htslib's reader"""
    import bam_build as bb

    d = data["d"]
    recs = []
    for i in range(0, data["n"], 2):
        cig = "150M" if i % 6 == 0 else "40M110S" if i % 6 == 2 else "20S100M5I25S"
        recs.append(bb.record(b"syn.%d" % i, flag=4 if i % 10 == 0 else 16 * (i % 4 == 0), mapq=60 if i % 14 else 3, cigar=cig))
    raw = bb.stream(recs)
    bp = _write(d / "aln.bam", bb.bgzf(raw))
    want = orc.clean_fastq(data["fq"][0], orc.set_from_bam(raw, 50, 0.5, 50), False)
    assert 0 < want.reads_out < want.reads_in
    o1, o2 = d / "b1.fq", d / "b2.fq"
    _run(cli, "alignment", "-i", data["r1"], "-o", o1, "-a", bp, "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50)
    assert open(o1, "rb").read() == want.written
    up = _write(d / "aln_uncompressed.dat", raw)  # `samtools view -u` style stream, explicit format
    _run(cli, "alignment", "-i", data["r1"], "-o", o2, "-a", up, "-f", "bam", "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50)
    assert open(o2, "rb").read() == want.written
    tp = _write(d / "trunc.bam", bb.bgzf(raw[:-7]))
    assert _run(cli, "alignment", "-i", data["r1"], "-o", d / "x.fq", "-a", tp, ok=False).returncode != 0
    cp = _write(d / "x.cram", b"CRAM\x03\x00" + b"\x00" * 40)
    assert _run(cli, "alignment", "-i", data["r1"], "-o", d / "x.fq", "-a", cp, ok=False).returncode != 0
