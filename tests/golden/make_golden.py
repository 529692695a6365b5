"""Writes tests/golden/vectors.json.

The reference (esteinig/scrubby 1.0.2) holds no tests or fixtures for this path and
cannot be built or imported here (Rust, un-vendored crates), so these vectors are
HAND-DERIVED from the cited reference lines: every `expect`
below was worked out by reading the reference code, not produced by running an
oracle.  tests/test_oracle_golden.py then checks both oracles against them.

Byte strings are stored latin-1 encoded so the JSON stays readable.

    python tests/golden/make_golden.py
"""
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def L(b: bytes) -> str:
    return b.decode("latin-1")


# --- 1. taxon state machine: classifier.rs:124-252, 345-373 -------------------------
REPORT_ROWS = [  # (rank code, taxid, name, direct reads)
    ("U", 0, "unclassified", 200), ("R", 1, "root", 5), ("R1", 131567, "cellular organisms", 3),
    ("D", 2759, "Eukaryota", 2), ("D1", 33154, "Opisthokonta", 1), ("K", 33208, "Metazoa", 4),
    ("K1", 6072, "Eumetazoa", 0), ("K2", 33213, "Bilateria", 6), ("K3", 33511, "Deuterostomia", 7),
    ("P", 7711, "Chordata", 8), ("P1", 89593, "Craniata", 9), ("C", 40674, "Mammalia", 10),
    ("O", 9443, "Primates", 0), ("F", 9604, "Hominidae", 11), ("G", 9605, "Homo", 12),
    ("S", 9606, "Homo sapiens", 430), ("C", 8782, "Aves", 10), ("K3", 33317, "Protostomia", 13),
    ("P", 6656, "Arthropoda", 14), ("C", 50557, "Insecta", 30), ("D", 2, "Bacteria", 15),
    ("S", 562, "Escherichia coli", 180),
]
METABULI_RANK = {"U": "no rank", "R": "no rank", "R1": "no rank", "D": "superkingdom", "D1": "clade",
                 "K1": "clade", "K2": "clade", "K3": "clade", "K": "kingdom", "P": "phylum",
                 "P1": "subphylum", "C": "class", "O": "order", "F": "family", "G": "genus",
                 "S": "species"}


def report(rows, full_word=False) -> bytes:
    out = []
    depth = 0
    for code, tid, name, direct in rows:
        rank = METABULI_RANK[code] if full_word else code
        # kraken2 style: percentage padded, clade count >= direct, name indented
        out.append(f"{1.5:6.2f}\t{direct + 7}\t{direct}\t{rank}\t{tid}\t{'  ' * depth}{name}")
        depth = min(depth + 1, 6)
    return ("\n".join(out) + "\n").encode()


CHORDATA = ["7711", "89593", "40674", "9604", "9605", "9606", "8782", "33317"]
taxon_cases = [
    dict(name="chordata_direct_9606", taxa=["Chordata"], direct=["9606"], expect=CHORDATA,
         why="K3 Protostomia has len 2 so it cannot close the subtree (classifier.rs:200); Primates has 0 direct reads"),
    dict(name="taxid_7711", taxa=["7711"], direct=[], expect=CHORDATA, why="taxa match by taxid string"),
    dict(name="metazoa", taxa=["Metazoa"], direct=[],
         expect=["33208", "33213", "33511", "7711", "89593", "40674", "9604", "9605", "9606", "8782",
                 "33317", "6656", "50557"], why="closed by 'D Bacteria' (Domain <= Kingdom, 1 char)"),
    dict(name="direct_only", taxa=[], direct=["9606"], expect=["9606"], why="taxa_direct inserts regardless of reads"),
    dict(name="direct_root", taxa=[], direct=["root"], expect=["1"], why="direct test precedes the <Domain skip"),
    dict(name="taxa_root", taxa=["root"], direct=[], expect=[], why="level < Domain is skipped before the taxa test (:157-166)"),
    dict(name="primates", taxa=["Primates"], direct=[], expect=["9604", "9605", "9606"],
         why="Primates itself has 0 direct reads; closed by 'C Aves'"),
    dict(name="trimmed_args", taxa=["  Chordata "], direct=[" 9606"], expect=CHORDATA, why="classifier.rs:132-133 trim"),
]
taxon_metabuli = dict(
    name="metabuli_fullword_never_closes", taxa=["Chordata"], direct=[],
    expect=["7711", "89593", "40674", "9604", "9605", "9606", "8782", "33317", "6656", "50557", "2", "562"],
    why="rank strings are never 1 char long, so the subtree is never closed (classifier.rs:200)")

# --- 2. PAF predicate: alignment.rs:100-108, 244-275 ---------------------------------


def paf(q, qlen, qs, qe, mapq, extra=""):
    return f"{q}\t{qlen}\t{qs}\t{qe}\t+\tchr1\t1000000\t100\t250\t140\t150\t{mapq}{extra}"


PAF_LINES = [
    paf("r1", 150, 0, 150, 60), paf("r2", 150, 0, 40, 60), paf("r3", 60, 10, 45, 60),
    paf("r4", 150, 0, 150, 49), paf("r5", 150, 0, 40, 60), paf("r5", 150, 0, 100, 10),
    paf("r6", 150, 0, 40, 60), paf("r6", 150, 0, 100, 60), paf("r7", 80, 0, 40, 50),
    paf("r8", 0, 0, 0, 60, "\ttp:A:P\tcm:i:10"),
]
PAF_BUF = ("\n".join(PAF_LINES) + "\n").encode()
paf_cases = [
    dict(name="l50_c05_q50", buf=L(PAF_BUF), min_len=50, min_cov=0.5, min_mapq=50,
         expect=["r1", "r3", "r6", "r7"],
         why="OR of len/cov per record AND mapq; r5 never passes on a single record; r7 exact >= on cov and mapq; r8 qlen 0 => cov 0"),
    dict(name="defaults_all_in", buf=L(PAF_BUF), min_len=0, min_cov=0.0, min_mapq=0,
         expect=["r1", "r2", "r3", "r4", "r5", "r6", "r7", "r8"], why="0/0/0 passes every line"),
    dict(name="plus_prefixed_int", buf=L(paf("p", "+150", 0, 150, 60).encode()), min_len=50, min_cov=0.5,
         min_mapq=50, expect=["p"], why="Rust from_str accepts one leading '+'; last line needs no newline"),
    dict(name="crlf_lines", buf=L((paf("a", 150, 0, 150, 60) + "\r\n" + paf("b", 150, 0, 150, 60) + "\r\n").encode()),
         min_len=50, min_cov=0.5, min_mapq=50, expect=["a", "b"], why="BufRead::lines strips CRLF; mapq is last column"),
    dict(name="qend_lt_qstart_wraps", buf=L(paf("w", 150, 100, 50, 60).encode() + b"\n"), min_len=50, min_cov=2.0,
         min_mapq=0, expect=["w"], why="usize subtraction wraps in release builds => huge alen (alignment.rs:265-267)"),
    dict(name="qname_verbatim", buf=L(paf(" q x ", 150, 0, 150, 60).encode() + b"\n"), min_len=0, min_cov=0.0,
         min_mapq=0, expect=[" q x "], why="column 1 inserted without trim"),
]
paf_errors = [
    dict(name="sci_notation", buf=L(paf("e", "1e3", 0, 150, 60).encode()), error=10),
    dict(name="empty_int", buf=L(paf("e", "", 0, 150, 60).encode()), error=10),
    dict(name="space_int", buf=L(paf("e", " 150", 0, 150, 60).encode()), error=10),
    dict(name="minus_int", buf=L(paf("e", "-1", 0, 150, 60).encode()), error=10),
    dict(name="mapq_256", buf=L(paf("e", 150, 0, 150, 256).encode()), error=10),
    dict(name="u64_overflow", buf=L(paf("e", "18446744073709551616", 0, 150, 60).encode()), error=10),
    dict(name="eleven_columns", buf=L(b"e\t150\t0\t150\t+\tchr1\t1000\t0\t150\t140\t150"), error=11),
    dict(name="blank_line", buf=L(paf("a", 150, 0, 150, 60).encode() + b"\n\n"), error=11, error_line=1),
    dict(name="bad_int_before_missing_col", buf=L(b"e\tx\t0"), error=10,
         why="fields are evaluated in order: qlen parse fails before fields[3] is indexed"),
    dict(name="invalid_utf8", buf=L(b"\xff" + paf("e", 150, 0, 150, 60).encode()), error=1),
]
PAF_MAX_U64 = dict(name="u64_max_ok", buf=L(paf("m", "18446744073709551615", 0, 150, 255).encode()),
                   min_len=150, min_cov=9.0, min_mapq=255, expect=["m"])
paf_cases.append(PAF_MAX_U64)

# --- 3. FASTQ framing / normalisation: needletail 0.5.1 + cleaner.rs:742-754 -------
fastq_cases = [
    dict(name="normalise_unix", buf=L(b"@a x\nAC\n+a x\nII\n@b\r\nGT\r\n+\r\nII"), ids=[], reverse=False,
         written=L(b"@a x\nAC\n+\nII\n@b\nGT\n+\nII\n"), reads_in=2, reads_out=2,
         why="separator text dropped, CR trimmed, final newline added; first line ending LF => Unix"),
    dict(name="normalise_windows", buf=L(b"@a x\r\nAC\r\n+\r\nII\r\n@b\nGT\n+b\nII\n"), ids=[], reverse=False,
         written=L(b"@a x\r\nAC\r\n+\r\nII\r\n@b\r\nGT\r\n+\r\nII\r\n"), reads_in=2, reads_out=2,
         why="first record CRLF => every record written with CRLF"),
    dict(name="leading_space_id", buf=L(b"@ a b\nAC\n+\nII\n@c\nGT\n+\nII\n"), ids=["a"], reverse=False,
         written=L(b"@c\nGT\n+\nII\n"), reads_in=2, reads_out=1, why="split_whitespace skips leading blanks: id 'a'"),
    dict(name="qual_starts_with_at_and_plus", buf=L(b"@r1\nACGT\n+\n@III\n@r2\nACGT\n+\n+III\n@r3\nAC\n+\n@+\n"),
         ids=["r2"], reverse=False, written=L(b"@r1\nACGT\n+\n@III\n@r3\nAC\n+\n@+\n"), reads_in=3, reads_out=2,
         why="framing is strictly line-mod-4"),
    dict(name="extract_mode", buf=L(b"@r1\nA\n+\nI\n@r2\nC\n+\nI\n@r3\nG\n+\nI\n"), ids=["r2", "zz"], reverse=True,
         written=L(b"@r2\nC\n+\nI\n"), reads_in=3, reads_out=1, why="cleaner.rs:751-753"),
    dict(name="two_trailing_blank_lines", buf=L(b"@r1\nA\n+\nI\n\n\n"), ids=[], reverse=False,
         written=L(b"@r1\nA\n+\nI\n"), reads_in=1, reads_out=1, why="check_end tolerates a blank tail"),
    dict(name="slash_suffix_no_match", buf=L(b"@r1/1 desc\nA\n+\nI\n"), ids=["r1"], reverse=False,
         written=L(b"@r1/1 desc\nA\n+\nI\n"), reads_in=1, reads_out=1, why="no /1 handling anywhere"),
    dict(name="tab_and_vt_delimit_id", buf=L(b"@r1\tx\nA\n+\nI\n@r2\x0bx\nA\n+\nI\n@r3\x1cx\nA\n+\nI\n"),
         ids=["r1", "r2", "r3"], reverse=False, written=L(b"@r3\x1cx\nA\n+\nI\n"), reads_in=3, reads_out=1,
         why="U+0009 and U+000B are White_Space, U+001C is not (python str.split trap)"),
    dict(name="nbsp_delimits_id", buf=L("@r1\u00a0x\nA\n+\nI\n@r2\u3000y\nA\n+\nI\n".encode("utf-8")),
         ids=["r1", "r2"], reverse=False, written=L(b""), reads_in=2, reads_out=0,
         why="U+00A0 and U+3000 are White_Space"),
    dict(name="interior_cr_kept", buf=L(b"@r1\rx\nA\n+\nI\n"), ids=["r1"], reverse=False, written=L(b""),
         reads_in=1, reads_out=0, why="CR is whitespace for get_id; only a trailing CR is trimmed from lines"),
    dict(name="empty_seq_record", buf=L(b"@r1\n\n+\n\n@r2\nA\n+\nI\n"), ids=["r2"], reverse=False,
         written=L(b"@r1\n\n+\n\n"), reads_in=2, reads_out=1, why="0-length sequence and quality are equal lengths"),
    dict(name="too_short_is_empty", buf=L(b"@a\nA"), ids=[], reverse=False, written=L(b""), reads_in=0, reads_out=0,
         empty_input=True, why="niffler needs 5 bytes: FileTooShort => treated as empty (utils.rs:365)"),
]
fastq_errors = [
    dict(name="header_only_at", buf=L(b"@\nAC\n+\nII\n"), error=9, error_record=0),
    dict(name="header_all_space", buf=L(b"@r0\nA\n+\nI\n@ \t\nAC\n+\nII\n"), error=9, error_record=1),
    dict(name="invalid_start", buf=L(b"@r0\nA\n+\nI\nr1\nAC\n+\nII\n"), error=3, error_record=1),
    dict(name="invalid_separator", buf=L(b"@r0\nA\n-\nI\n"), error=4, error_record=0),
    dict(name="unequal_lengths", buf=L(b"@r0\nACG\n+\nII\n"), error=5, error_record=0),
    dict(name="unequal_after_cr_trim", buf=L(b"@r0\nAC\r\n+\nII\r\r\n"), error=5, error_record=0),
    dict(name="truncated", buf=L(b"@r0\nA\n+\nI\n@r1\nAC\n"), error=6, error_record=1),
    dict(name="three_blank_lines", buf=L(b"@r0\nA\n+\nI\n\n\n\n"), error=3, error_record=1,
         why="3 newlines reach SearchPosition::Quality, then validate sees '\\n' as start byte"),
    dict(name="blank_between_records", buf=L(b"@r0\nA\n+\nI\n\n@r1\nA\n+\nI\n"), error=3, error_record=1),
    dict(name="unknown_format", buf=L(b"hello world\n"), error=7, error_record=0),
    dict(name="invalid_utf8_header", buf=L(b"@r0 \xff\nA\n+\nI\n"), error=8, error_record=0),
    dict(name="overlong_utf8_header", buf=L(b"@r0 \xc0\xaf\nA\n+\nI\n"), error=8, error_record=0),
]

# --- 4. diff / report: utils.rs:250-285, report.rs:24-57 -----------------------------


def fq(ids):
    return "".join(f"@{i} d\nACGT\n+\nIIII\n" for i in ids).encode()


diff_cases = [
    dict(name="deplete_b_c", pairs=[[L(fq("abcc")), L(fq("a"))], [L(fq("abcc")), L(fq("a"))]],
         reads_in=8, reads_out=2, difference=6, diff_ids=["b", "c"],
         why="mates with the same id count twice but appear once in the id list"),
    dict(name="extract_b_c", pairs=[[L(fq("abcc")), L(fq("bcc"))], [L(fq("abcc")), L(fq("bcc"))]],
         reads_in=8, reads_out=6, difference=2, diff_ids=["a"],
         why="with -e the difference is the reads NOT written (report.rs:44-45)"),
    dict(name="empty_output", pairs=[[L(fq("ab")), L(b"")]], reads_in=2, reads_out=0, difference=2, diff_ids=["a", "b"]),
]
report_case = dict(  # README.md:198-200
    reads_in=6678, reads_out=3346, reads_removed=3332, reads_extracted=0)

# --- 5. id matching / reads files: classifier.rs:270-328, 401-419 --------------------
reads_cases = [
    dict(name="kraken_basic", style=0, taxids=["9606", "7711"],
         buf=L(b"C\tr1\t9606\t150|150\t9606:5\nU\tr2\t0\t150\t0:1\nC\t r3 \t 7711 \t150\tx\nC\tr4\t09606\t150\tx\n"
               b"C\tr5\t+9606\t150\tx\nC\tr6\tHomo sapiens (taxid 9606)\t150\tx\n"),
         expect=["r1", "r3"], why="fields trimmed; taxid compared as a string, so 09606/+9606/names do not match"),
    dict(name="metabuli_basic", style=1, taxids=["9606"],
         buf=L(b"1\tread1\t9606\t100\t80.5\tspecies\tann\n0\tread2\t0\t100\t0\tno rank\t-\n"), expect=["read1"]),
    dict(name="kraken_crlf_no_final_newline", style=0, taxids=["1"], buf=L(b"C\ta\t1\t1\tx\r\nC\tb\t1\t1\tx"),
         expect=["a", "b"]),
    dict(name="nonnumeric_taxid_string", style=0, taxids=["A12"], buf=L(b"C\ta\tA12\t1\tx\nC\tb\tA1\t1\tx\n"),
         expect=["a"], why="membership is string equality on whatever the report held"),
]
reads_errors = [
    dict(name="kraken_four_columns", style=0, buf=L(b"C\tr1\t9606\t150\n"), error=11),
    dict(name="metabuli_six_columns", style=1, buf=L(b"1\tr1\t9606\t100\t80.5\tgenus\n"), error=11),
]
txt_cases = [
    dict(name="verbatim_lines", buf=L(b"r1\nr2 \n\n r3\r\nr4"), expect=["", " r3", "r1", "r2 ", "r4"],
         why="alignment.rs:72-75: no trim, no split; blank line is the empty id"),
]
get_id_cases = [
    dict(header=L(b"read1 description"), expect="read1"),
    dict(header=L(b"  a b"), expect="a"),
    dict(header=L(b"syn.12 1:N:0:ATCACG"), expect="syn.12"),
    dict(header=L("id\u2003rest".encode("utf-8")), expect="id", why="U+2003 EM SPACE is White_Space"),
    dict(header=L("id\u200bsame".encode("utf-8")), expect=L("id\u200bsame".encode("utf-8")),
         why="U+200B ZERO WIDTH SPACE is not White_Space"),
    dict(header=L(b"id\x85x"), error=8, why="lone 0x85 is invalid UTF-8 (NEL is C2 85)"),
    dict(header=L(b"id\xc2\x85x"), expect="id"),
]

# --- 6. SAM evidence: alignment.rs:117-146, 154-211 (htslib feature), text SAM ---------------------
SAM_HDR = "@HD\tVN:1.6\tSO:unsorted\n@SQ\tSN:chr1\tLN:100000\n@PG\tID:minimap2\n"


def sam(q, flag, rname, pos, mapq, cigar, seqlen, qual=None, extra=""):
    seq = "*" if seqlen is None else "ACGT" * (seqlen // 4) + "ACGT"[: seqlen % 4]
    if qual is None:
        qual = "*" if seqlen is None else "I" * seqlen
    return f"{q}\t{flag}\t{rname}\t{pos}\t{mapq}\t{cigar}\t*\t0\t0\t{seq}\t{qual}{extra}"


SAM_BASE = SAM_HDR + "\n".join([
    sam("s1", 0, "chr1", 100, 60, "150M", 150),
    sam("s2", 0, "chr1", 100, 60, "40M110S", 150),            # qalen 40, cov 0.267
    sam("s3", 16, "chr1", 100, 60, "10S35M15S", 60),          # qalen 35, cov 0.583 rescues
    sam("s4", 0, "chr1", 100, 49, "150M", 150),               # mapq below
    sam("s5", 4, "*", 0, 0, "*", 150),                        # unmapped flag
    sam("s6", 0, "chr1", 100, 60, "50M50I50M", 150),          # insertions count: qalen 150
    sam("s7", 0, "chr1", 100, 60, "100=50M", 150),            # '=' is not Cigar::Match: qalen 50
    sam("s8", 0, "chr1", 100, 60, "101X49M", 150),            # qalen 49, cov 0.327
    sam("s9", 0, "chr1", 100, 60, "50M100D50M", 100),         # deletions do not count: qalen 100
    sam("s10", 16, "chr1", 0, 60, "150M", 150),               # POS 0: htslib treats it as unmapped
    sam("s11", 0, "chr1", 100, 60, "150M", None),             # SEQ '*': qlen 0, cov 0, qalen 150 >= 50
    sam("s12", "0x10", "chr1", 100, 60, "150M", 150),         # strtol(.., 0): hex flag
    sam("s13", "0x4", "chr1", 100, 60, "150M", 150),          # hex unmapped
    sam("s14", 256, "chr1", 100, 10, "150M", 150),            # same read: secondary fails ...
    sam("s14", 0, "chr1", 900, 60, "20H130M", 130, extra="\tNM:i:0\tAS:i:130"),  # ... primary passes (qalen 130)
    sam("s15", 0, "*", 100, 60, "150M", 150),                 # RNAME '*': unmapped
    sam("s16", 0, "chr1", 100, 50, "25M25N25M50S", 100),      # qalen 50 == min_len, mapq == min_mapq
]) + "\n"
sam_cases = [
    dict(name="predicate_50_0.5_50", buf=L(SAM_BASE.encode()), min_len=50, min_cov=0.5, min_mapq=50,
         expect=["s1", "s3", "s6", "s7", "s9", "s11", "s12", "s14", "s16"],
         why="M and I only (alignment.rs:160-168), OR of len/cov, >= comparisons, unmapped skipped"),
    dict(name="defaults_all_mapped", buf=L(SAM_BASE.encode()), min_len=0, min_cov=0.0, min_mapq=0,
         expect=["s1", "s2", "s3", "s4", "s6", "s7", "s8", "s9", "s11", "s12", "s14", "s16"],
         why="every mapped record passes; s5 s10 s13 s15 are unmapped"),
    dict(name="crlf_and_no_final_newline", buf=L((SAM_HDR.replace("\n", "\r\n") + sam("a", 0, "chr1", 5, 60, "150M", 150)
                                                + "\r\n" + sam("b", 0, "chr1", 5, 60, "10M", 10)).encode()),
         min_len=50, min_cov=0.5, min_mapq=0, expect=["a", "b"], why="b: qalen 10 < 50 but cov 1.0"),
    dict(name="header_only", buf=L(SAM_HDR.encode()), min_len=0, min_cov=0.0, min_mapq=0, expect=[]),
    dict(name="octal_flag", buf=L((SAM_HDR + sam("o", "04", "chr1", 5, 60, "150M", 150) + "\n"
                                  + sam("p", "020", "chr1", 5, 60, "150M", 150) + "\n").encode()),
         min_len=0, min_cov=0.0, min_mapq=0, expect=["p"], why="04 = unmapped; 020 = 16 reverse strand"),
    dict(name="undeclared_rname_is_unmapped",
         buf=L((SAM_HDR + sam("k", 0, "chr1", 5, 60, "150M", 150) + "\n" + sam("u", 0, "chr2", 5, 60, "150M", 150) + "\n"
                + sam("v", 0, "chr", 5, 60, "150M", 150) + "\n").encode()),
         min_len=0, min_cov=0.0, min_mapq=0, expect=["k"],
         why="htslib sam_parse1: 'unrecognized reference name; treated as unmapped' (tid -1 -> BAM_FUNMAP), "
             "rust-htslib is_unmapped() skips it (alignment.rs:132-134)"),
    dict(name="two_sq_lines", buf=L((SAM_HDR + "@SQ\tLN:5\tSN:chr2\n" + sam("u", 0, "chr2", 5, 60, "150M", 150) + "\n").encode()),
         min_len=0, min_cov=0.0, min_mapq=0, expect=["u"], why="SN: may be any field of the @SQ line"),
    dict(name="no_header_only_unmapped", buf=L((sam("s5", 4, "*", 0, 0, "*", 150) + "\n").encode()),
         min_len=0, min_cov=0.0, min_mapq=0, expect=[], why="RNAME '*' needs no header"),
]
sam_errors = [
    dict(name="ten_fields", buf=L(("\t".join(sam("e", 0, "chr1", 5, 60, "150M", 150).split("\t")[:10]) + "\n").encode()), error=22),
    dict(name="mapq_256", buf=L((sam("e", 0, "chr1", 5, 256, "150M", 150) + "\n").encode()), error=22),
    dict(name="bad_cigar_op", buf=L((sam("e", 0, "chr1", 5, 60, "150Q", 150) + "\n").encode()), error=22),
    dict(name="cigar_without_count", buf=L((sam("e", 0, "chr1", 5, 60, "M", 150) + "\n").encode()), error=22),
    dict(name="cigar_seq_mismatch", buf=L((sam("e", 0, "chr1", 5, 60, "149M", 150) + "\n").encode()), error=22),
    dict(name="qual_len_mismatch", buf=L((sam("e", 0, "chr1", 5, 60, "150M", 150, qual="I" * 149) + "\n").encode()), error=22),
    dict(name="flag_text", buf=L((sam("e", "abc", "chr1", 5, 60, "150M", 150) + "\n").encode()), error=22),
    dict(name="blank_line", buf=L((SAM_HDR + sam("a", 0, "chr1", 5, 60, "150M", 150) + "\n\n"
                                  + sam("b", 0, "chr1", 5, 60, "150M", 150) + "\n").encode()), error=22, error_line=4),
    dict(name="error_after_header", buf=L((SAM_HDR + sam("a", 0, "chr1", 5, 60, "150M", 150) + "\n"
                                          + sam("e", 0, "chr1", "x", 60, "150M", 150) + "\n").encode()), error=22, error_line=4),
    dict(name="qname_not_utf8", buf=L(SAM_HDR.encode() + b"\xff" + (sam("e", 0, "chr1", 5, 60, "150M", 150) + "\n").encode()),
         error=8, error_line=3),
    dict(name="no_sq_lines", buf=L((sam("a", 4, "*", 0, 0, "*", 150) + "\n" + sam("b", 0, "chr1", 5, 60, "150M", 150) + "\n").encode()),
         error=22, error_line=1, why="htslib sam_parse1: 'no SQ lines present in the header' for any RNAME other than '*'"),
]

# --- FASTA input: needletail 0.5.1 fasta reader + write_fasta under cleaner.rs:742-754 (hand-derived) -------------
fasta_cases = [
    dict(name="multi_line_verbatim", buf=L(b">a x\nAC\nGT\n>b\nTT\n>c\n>d\nAA"), ids=["b"], reverse=False,
         written=L(b">a x\nAC\nGT\n>c\n\n>d\nAA\n"), other=L(b">b\nTT\n")),
    dict(name="extract", buf=L(b">a x\nAC\nGT\n>b\nTT\n>c\n>d\nAA"), ids=["b"], reverse=True,
         written=L(b">b\nTT\n"), other=L(b">a x\nAC\nGT\n>c\n\n>d\nAA\n")),
    dict(name="crlf_file", buf=L(b">a\r\nAC\r\nGT\r\n>b\r\nTT\r\n"), ids=["b"], reverse=False,
         written=L(b">a\r\nAC\r\nGT\r\n"), other=L(b">b\r\nTT\r\n")),
    dict(name="blank_last_line", buf=L(b">a\nAC\n\n"), ids=[], reverse=False, written=L(b">a\nAC\n\n"), other=""),
    dict(name="line_ending_from_first_record_with_a_newline", buf=L(b">a\n>b\r\nAC\r\n"), ids=[], reverse=False,
         written=L(b">a\n\n>b\r\nAC\r\n"), other=""),
    dict(name="one_trailing_cr_trimmed", buf=L(b">a\nAC\r\n>b\n\nTT\r"), ids=["a"], reverse=True,
         written=L(b">a\nAC\n"), other=L(b">b\n\nTT\n")),
    dict(name="id_is_first_token", buf=L(b">  a b\nAC\n"), ids=["a"], reverse=False, written="", other=L(b">  a b\nAC\n")),
]
fasta_errors = [
    dict(name="header_without_newline", buf=L(b">abcd"), error=6, error_record=0),
    dict(name="last_header_without_sequence_line", buf=L(b">a\nAC\n>b\n"), error=6, error_record=1),
    dict(name="empty_id", buf=L(b">a\nAC\n>\nTT\n"), error=9, error_record=1),
    dict(name="id_not_utf8", buf=L(b">a\nAC\n>\xff\nTT\n"), error=8, error_record=1),
]

doc = dict(
    provenance="hand-derived from /root/reference/src (see make_golden.py); reference has no fixtures; parity unpinned",
    report_kraken=L(report(REPORT_ROWS)), report_metabuli=L(report(REPORT_ROWS, True)),
    taxon_cases=taxon_cases, taxon_metabuli=taxon_metabuli, paf_cases=paf_cases, paf_errors=paf_errors,
    fastq_cases=fastq_cases, fastq_errors=fastq_errors, diff_cases=diff_cases, report_case=report_case,
    reads_cases=reads_cases, reads_errors=reads_errors, txt_cases=txt_cases, get_id_cases=get_id_cases,
    sam_cases=sam_cases, sam_errors=sam_errors, fasta_cases=fasta_cases, fasta_errors=fasta_errors)

if __name__ == "__main__":
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(doc, f, indent=1, ensure_ascii=True)
    print("wrote", os.path.join(HERE, "vectors.json"))
