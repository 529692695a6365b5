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


def _run(cli, *args, ok=True, env=None):
    r = subprocess.run([cli, *map(str, args)], capture_output=True, text=True, timeout=300, env=env)
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


# ---------------------------------------------------------------------------- `scrubby reads` with stand-in tools
FAKE_TOOL = r"""#!/bin/sh
# stand-in for an external tool (TEST ONLY): logs its command line, then plays back prepared outputs
name=$(basename "$0")
echo "$name $*" >> "$FAKE_LOG"
case "$name $1" in
  "kraken2 --version"|"minigraph --version"|"minimap2 --version"|"bowtie2 --version"|"strobealign --version") exit 0;;
esac
case "$name" in
  kraken2)
    while [ $# -gt 0 ]; do
      case "$1" in --output) cp "$FAKE_READS" "$2"; shift;; --report) cp "$FAKE_REPORT" "$2"; shift;; esac
      shift
    done;;
  metabuli)
    [ $# -eq 0 ] && exit 0
    for a in "$@"; do prev2=$prev; prev=$a; done   # ... <index> <dir> metabuli
    cp "$FAKE_REPORT" "$prev2/metabuli_report.tsv"; cp "$FAKE_READS" "$prev2/metabuli_classifications.tsv";;
  minigraph) cat "$FAKE_PAF"; exit ${FAKE_EXIT:-0};;
  minimap2|bowtie2|strobealign) echo "@HD	VN:1.6";;
  samtools)
    if [ "$1" = view ]; then cat; else
      for a in "$@"; do case "$prev" in -0|-1) cat > "$a";; -2) : > "$a";; esac; prev=$a; done
    fi;;
esac
"""


@pytest.fixture(scope="module")
def fake_tools(data):
    d = data["d"] / "bin"
    d.mkdir()
    for name in ("kraken2", "metabuli", "minigraph", "minimap2", "bowtie2", "strobealign", "samtools"):
        p = d / name
        p.write_text(FAKE_TOOL)
        p.chmod(0o755)
    idx = data["d"] / "k2db"
    idx.mkdir()
    mmi = _write(data["d"] / "ref.mmi", b"x")
    # Metabuli-style outputs: full-word ranks in the report, seven columns per read
    mb_reads = b"".join(l + b"\t0.9\tspecies\n" for l in data["kr"].split(b"\n") if l)
    env = dict(os.environ, PATH=f"{d}:{os.environ['PATH']}", FAKE_LOG=str(data["d"] / "tools.log"),
               FAKE_READS=data["kr_p"], FAKE_REPORT=data["rep_p"], FAKE_PAF=data["paf_p"])
    return dict(env=env, idx=str(idx), mmi=mmi, log=data["d"] / "tools.log", mb_reads=_write(data["d"] / "mb.tsv", mb_reads))


def test_cli_reads_kraken2_and_metabuli(cli, data, fake_tools):
    d, env = data["d"], fake_tools["env"]
    tax = orc.taxids_from_report(data["rep"], ["Chordata"], ["9606"])
    want = [orc.clean_fastq(f, orc.set_from_reads(data["kr"], 0, tax), False) for f in data["fq"]]
    o1, o2, js = d / "k1.fq", d / "k2.fq", d / "k.json"
    open(fake_tools["log"], "w").close()
    _run(cli, "reads", "-i", data["r1"], data["r2"], "-o", o1, o2, "-I", fake_tools["idx"], "-c", "kraken2", "-T", "Chordata",
         "-D", "9606", "-t", 7, "-C", "--confidence 0.1", "-w", d / "work", "-j", js, env=env)
    assert open(o1, "rb").read() == want[0].written and open(o2, "rb").read() == want[1].written
    log = open(fake_tools["log"]).read().split("\n")
    assert log[0] == "kraken2 --version"                       # cleaner.rs:277-284
    assert log[1] == (f"kraken2 --threads 7 --db {fake_tools['idx']} --confidence 0.1 --paired {data['r1']} {data['r2']} "
                      f"--output {d / 'work' / 'kraken.reads'} --report {d / 'work' / 'kraken.report'}")  # cleaner.rs:300-311
    rep = json.load(open(js))
    assert rep["settings"]["classifier"] == "kraken2" and rep["settings"]["classifier_args"] == "--confidence 0.1"
    assert rep["reads_in"] == 2 * data["n"] and rep["reads_out"] == want[0].reads_out + want[1].reads_out
    # metabuli, single-end, extract
    env2 = dict(env, FAKE_READS=fake_tools["mb_reads"])
    o3 = d / "m1.fq"
    open(fake_tools["log"], "w").close()
    _run(cli, "reads", "-i", data["r1"], "-o", o3, "-I", fake_tools["idx"], "-c", "metabuli", "-T", "Chordata", "-D", "9606",
         "-w", d / "work", "-e", env=env2)
    mb = open(fake_tools["mb_reads"], "rb").read()
    want_m = orc.clean_fastq(data["fq"][0], orc.set_from_reads(mb, 1, tax), True)
    assert open(o3, "rb").read() == want_m.written and want_m.reads_out > 0
    log = open(fake_tools["log"]).read().split("\n")
    assert log[0] == "metabuli " and log[1] == (f"metabuli classify --seq-mode 3 --threads 4 {data['r1']} {fake_tools['idx']} "
                                                f"{d / 'work'} metabuli")  # cleaner.rs:352-361
    # validation (scrubby.rs:813-867): classifier without taxa; index that is not a directory; both tools at once
    for extra in (["-c", "kraken2"], ["-c", "kraken2", "-T", "x", "-I", data["r1"]], ["-c", "kraken2", "-a", "minimap2", "-T", "x"]):
        args = ["reads", "-i", data["r1"], "-o", d / "x.fq", "-I", fake_tools["idx"]] + extra
        assert _run(cli, *args, ok=False, env=env).returncode != 0
    # the tool is probed before anything runs (cleaner.rs:110-123)
    bare = dict(env, PATH="/usr/bin:/bin")
    r = _run(cli, "reads", "-i", data["r1"], "-o", d / "x.fq", "-I", fake_tools["idx"], "-c", "kraken2", "-T", "x", ok=False, env=bare)
    assert "kraken2" in r.stderr


def test_cli_reads_aligners(cli, data, fake_tools):
    d, env = data["d"], fake_tools["env"]
    # minigraph: stdout PAF comes back into the GPU path (cleaner.rs:651-687); defaults -l 0 -c 0 -q 0: every line's qname
    want = orc.clean_fastq(data["fq"][0], orc.set_from_paf(data["paf"], 0, 0.0, 0), False)
    o1 = d / "mg.fq"
    open(fake_tools["log"], "w").close()
    _run(cli, "reads", "-i", data["r1"], "-o", o1, "-I", fake_tools["mmi"], "-a", "minigraph", "-A", "-k 15", env=env)
    assert open(o1, "rb").read() == want.written and 0 < want.reads_out < want.reads_in
    log = open(fake_tools["log"]).read().split("\n")
    assert log[:2] == ["minigraph --version", f"minigraph -cx lr -N 0 -t 4 -k 15 {fake_tools['mmi']} {data['r1']}"]  # single-end default: lr
    # a failing child is CommandFailed even when its PAF parsed (cleaner.rs:681-684)
    assert _run(cli, "reads", "-i", data["r1"], "-o", d / "x.fq", "-I", fake_tools["mmi"], "-a", "minigraph", ok=False,
                env=dict(env, FAKE_EXIT="3")).returncode != 0
    assert _run(cli, "reads", "-i", data["r1"], "-o", d / "x.fq", "-I", fake_tools["mmi"], "-a", "minigraph", "-p", "map-ont",
                ok=False, env=env).returncode != 0  # MinigraphPresetNotSupported
    # SAM-emitting aligners: the reference's shell pipeline, samtools does the depletion (cleaner.rs:46-71, 385-410)
    open(fake_tools["log"], "w").close()
    o2, o3 = d / "mm1.fq", d / "mm2.fq"
    _run(cli, "reads", "-i", data["r1"], data["r2"], "-o", o2, o3, "-I", fake_tools["mmi"], "-a", "minimap2", "-t", 3, env=env)
    log = open(fake_tools["log"]).read().split("\n")
    assert log[0] == "minimap2 --version"
    assert f"minimap2 -ax sr --secondary=no -t 3 {fake_tools['mmi']} {data['r1']} {data['r2']}" in log  # paired default: sr
    assert "samtools view -h -f 12 -" in log
    assert f"samtools fastq --threads 4 -s /dev/null -c 6 -n -1 {o2} -2 {o3}" in log
    assert os.path.exists(o2) and os.path.exists(o3)
    # defaults: single-end without -a / -c is minimap2 map-ont; minimap2 refuses the lr preset; bowtie2 wants its index files
    open(fake_tools["log"], "w").close()
    _run(cli, "reads", "-i", data["r1"], "-o", d / "mm3.fq", "-I", fake_tools["mmi"], "-e", env=env)
    log = open(fake_tools["log"]).read().split("\n")
    assert f"minimap2 -ax map-ont --secondary=no -t 4 {fake_tools['mmi']} {data['r1']}" in log and "samtools view -h -F 4 -" in log
    for extra in (["-a", "minimap2", "-p", "lr"], ["-a", "bowtie2"], ["-a", "strobealign", "-I", d / "missing.fa"]):
        assert _run(cli, "reads", "-i", data["r1"], "-o", d / "x.fq", "-I", fake_tools["mmi"], *extra, ok=False, env=env).returncode != 0
    for ext in ("1.bt2", "2.bt2", "3.bt2", "4.bt2", "rev.1.bt2", "rev.2.bt2"):
        _write(d / f"bt.{ext}", b"x")
    open(fake_tools["log"], "w").close()
    _run(cli, "reads", "-i", data["r1"], "-o", d / "bt.fq", "-I", d / "bt", "-a", "bowtie2", env=env)
    assert f"bowtie2 -x {d / 'bt'} -U {data['r1']} -k 1 --mm -p 4\n" in open(fake_tools["log"]).read()


def test_cli_fasta_reads(cli, data):
    """FASTA input through the CLI (needletail's FASTA reader under the same clean_reads loop): multi-line records,
    TXT id list, deplete and extract, report JSON from the diff over the FASTA files"""
    import random

    d = data["d"]
    rng = random.Random(4)
    recs = []
    for i in range(3000):
        lines = [bytes(rng.choice(b"ACGT") for _ in range(70)) for _ in range(1 + i % 4)]
        recs.append(b">read%d runid=x\n" % i + b"\n".join(lines) + b"\n")
    fa = b"".join(recs)
    ids = b"".join(b"read%d\n" % i for i in range(0, 3000, 4))
    fp, ip = _write(d / "reads.fasta", fa), _write(d / "ids.txt", ids)
    for extract in (False, True):
        o, js = d / f"fa_{extract}.fasta", d / f"fa_{extract}.json"
        args = ["alignment", "-i", fp, "-o", o, "-a", ip, "-j", js] + (["-e"] if extract else [])
        _run(cli, *args)
        want = orc.clean_fastq(fa, orc.set_from_txt(ids), extract)
        assert open(o, "rb").read() == want.written and 0 < want.reads_out < want.reads_in
        rep = json.load(open(js))
        assert (rep["reads_in"], rep["reads_out"]) == (3000, want.reads_out)
        assert (rep["reads_removed"], rep["reads_extracted"]) == ((0, 3000 - want.reads_out) if extract else (3000 - want.reads_out, 0))


def test_cli_tiny_gz_id_list_is_not_empty(cli, data, tmp_path):
    """ADVICE r01 / utils.rs:359-375: niffler's 5-byte FileTooShort rule applies to the RAW file; a .gz id list whose
    decompressed content is "syn.5\\n" (6 bytes) or even 3 bytes is read, only a raw file under 5 bytes is empty"""
    ids_gz = tmp_path / "ids.gz"
    with gzip.open(ids_gz, "wb") as f:
        f.write(b"s\n")  # 2 bytes decompressed: matches nothing, but is NOT "empty"
    o1 = tmp_path / "o1.fq"
    js = tmp_path / "r.json"
    _run(cli, "alignment", "-i", data["r1"], "-o", o1, "-a", ids_gz, "--format", "txt", "-j", js)
    assert open(o1, "rb").read() == data["fq"][0]
    one = tmp_path / "one.gz"
    with gzip.open(one, "wb") as f:
        f.write(b"syn.5\n")
    _run(cli, "alignment", "-i", data["r1"], "-o", o1, "-a", one, "--format", "txt", "-j", js)
    want = orc.clean_fastq(data["fq"][0], orc.OSet.from_ids([b"syn.5"])).written
    assert open(o1, "rb").read() == want and len(want) < len(data["fq"][0])
    raw3 = _write(tmp_path / "tiny.txt", b"s\n")  # raw file of 2 bytes: empty by the FileTooShort rule
    _run(cli, "alignment", "-i", data["r1"], "-o", o1, "-a", raw3, "--format", "txt", "-j", js)
    assert open(o1, "rb").read() == data["fq"][0]


@pytest.mark.parametrize("chunk,halo", [(300_000, 20_000), (1 << 20, 1 << 16), (4000, 700)])
def test_cli_gzip_input_streams_through_the_shard_entry_point(cli, data, chunk, halo):
    """a plain-gzip FASTQ goes inflate -> sgpu_clean_fastq_shard chunk by chunk -> deflate as a pipeline
    (scrubby_host.cpp: clean_fastq_gz_stream); small chunks force many shard calls, paired files run two pipelines at once.
    SCRUBBY_NO_GZ_STREAM takes the whole-file path: the same bytes either way."""
    d = data["d"]
    want = [orc.clean_fastq(f, orc.set_from_paf(data["paf"], 50, 0.5, 50), False).written for f in data["fq"]]
    n_in = 20_000 if chunk > 4000 else 600
    ins = []
    for m in (0, 1):
        g = d / f"s{m}_{chunk}.fq.gz"
        src = data["fq"][m] if chunk > 4000 else data["fq"][m][: data["fq"][m].index(b"\n@syn.%d " % n_in) + 1]
        with open(g, "wb") as f:  # two members, cut inside a record
            f.write(gzip.compress(src[: len(src) // 3], 1) + gzip.compress(src[len(src) // 3:], 1))
        ins.append(g)
    if chunk <= 4000:
        want = [orc.clean_fastq(gzip.open(g).read(), orc.set_from_paf(data["paf"], 50, 0.5, 50), False).written for g in ins]
    for stream in (True, False):
        env = dict(os.environ, SCRUBBY_STREAM_CHUNK=str(chunk), SCRUBBY_STREAM_HALO=str(halo))
        if not stream:
            env["SCRUBBY_NO_GZ_STREAM"] = "1"
        o = [str(d / f"so{m}_{chunk}_{stream}.fq") + (".gz" if m else "") for m in (0, 1)]
        js = d / f"s_{chunk}_{stream}.json"
        _run(cli, "alignment", "-i", *ins, "-o", *o, "-a", data["paf_p"], "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50,
             "-j", js, env=env)
        assert open(o[0], "rb").read() == want[0]
        assert gzip.open(o[1], "rb").read() == want[1]
        rep = json.load(open(js))
        assert rep["reads_in"] == 2 * n_in


def test_cli_gzip_stream_parse_error_and_crlf(cli, data, tmp_path):
    """a parse error in a later chunk of the stream: exit status and message as on the whole-file path, the records before
    it are in the output; a CRLF file keeps the first record's line ending decision across chunks"""
    fq = bytearray(data["fq"][0][: data["fq"][0].index(b"\n@syn.900 ") + 1])
    at = fq.index(b"\n+\n", fq.index(b"@syn.700 ")) + 1
    fq[at] = ord("-")
    bad = tmp_path / "bad.fq.gz"
    bad.write_bytes(gzip.compress(bytes(fq), 1))
    want = orc.clean_fastq(bytes(fq), orc.set_from_paf(data["paf"], 50, 0.5, 50), False, raise_on_error=False)
    assert want.error == 4 and want.error_record == 700
    outs = []
    for stream in (True, False):
        env = dict(os.environ, SCRUBBY_STREAM_CHUNK="50000", SCRUBBY_STREAM_HALO="4000")
        if not stream:
            env["SCRUBBY_NO_GZ_STREAM"] = "1"
        o = tmp_path / f"bad_{stream}.fq"
        r = _run(cli, "alignment", "-i", bad, "-o", o, "-a", data["paf_p"], "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50,
                 ok=False, env=env)
        outs.append((r.returncode, r.stderr.strip().split("\n")[-1], open(o, "rb").read()))
        assert outs[-1][2] == want.written
    assert outs[0] == outs[1]
    crlf = data["fq"][1][: data["fq"][1].index(b"\n@syn.1500 ") + 1].replace(b"\n", b"\r\n")
    src = tmp_path / "crlf.fq.gz"
    src.write_bytes(gzip.compress(crlf, 1))
    wantc = orc.clean_fastq(crlf, orc.set_from_paf(data["paf"], 50, 0.5, 50), False)
    assert wantc.crlf
    env = dict(os.environ, SCRUBBY_STREAM_CHUNK="70000", SCRUBBY_STREAM_HALO="5000")
    o = tmp_path / "crlf_out.fq"
    _run(cli, "alignment", "-i", src, "-o", o, "-a", data["paf_p"], "--min-len", 50, "--min-cov", 0.5, "--min-mapq", 50, env=env)
    assert open(o, "rb").read() == wantc.written
