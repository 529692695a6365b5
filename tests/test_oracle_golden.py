"""Both CPU oracles against the hand-derived golden vectors."""
import pytest

from conftest import B
from oracle import oracle as orc
from oracle import pyoracle as pyo


def _ids(lst):
    return sorted(B(x) for x in lst)


def test_error_codes_match_header():
    import re, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    def codes(path, prefix):
        txt = open(path).read()
        return {m.group(1): int(m.group(2)) for m in re.finditer(prefix + r"_(ERR_\w+|OK)\s*=\s*(\d+)", txt)}
    o = codes(os.path.join(root, "oracle", "scrubby_oracle.h"), "ORC")
    g = codes(os.path.join(root, "include", "scrubby_gpu.h"), "SGPU")
    assert o and all(g.get(k) == v for k, v in o.items()), (o, g)


def test_taxon_state_machine(golden):
    rep = B(golden["report_kraken"])
    for c in golden["taxon_cases"]:
        want = _ids(c["expect"])
        assert orc.taxids_from_report(rep, c["taxa"], c["direct"]).sorted_ids() == want, c["name"]
        assert sorted(pyo.taxids_from_report(rep, c["taxa"], c["direct"])) == want, c["name"]
    c = golden["taxon_metabuli"]
    rep = B(golden["report_metabuli"])
    assert orc.taxids_from_report(rep, c["taxa"], c["direct"]).sorted_ids() == _ids(c["expect"])
    assert sorted(pyo.taxids_from_report(rep, c["taxa"], c["direct"])) == _ids(c["expect"])


def test_paf_predicate(golden):
    for c in golden["paf_cases"]:
        want = _ids(c["expect"])
        got = orc.set_from_paf(B(c["buf"]), c["min_len"], c["min_cov"], c["min_mapq"]).sorted_ids()
        assert got == want, c["name"]
        assert sorted(pyo.ids_from_paf(B(c["buf"]), c["min_len"], c["min_cov"], c["min_mapq"])) == want, c["name"]


def test_paf_errors(golden):
    for c in golden["paf_errors"]:
        with pytest.raises(orc.OracleError) as e:
            orc.set_from_paf(B(c["buf"]), 0, 0.0, 0)
        assert e.value.code == c["error"], c["name"]
        if "error_line" in c:
            assert e.value.index == c["error_line"]
        with pytest.raises(pyo.RefError) as e2:
            pyo.ids_from_paf(B(c["buf"]))
        assert e2.value.code == c["error"], c["name"]


def _random_sam(seed: int, n: int = 400) -> bytes:
    """well-formed SAM lines with every CIGAR operator, hex / octal flags, '*' fields and CRLF endings"""
    import random

    rnd = random.Random(seed)
    out = ["@HD\tVN:1.6", "@SQ\tSN:chr1\tLN:1000000", "@SQ\tLN:5000\tSN:chrM"]
    for i in range(n):
        ops, qlen = [], 0
        if rnd.random() < 0.9:
            for _ in range(rnd.randint(1, 6)):
                op = rnd.choice("MMMIDNSHP=X")
                c = rnd.randint(1, 90)
                ops.append(f"{c}{op}")
                if op in "MIS=X":
                    qlen += c
        cigar = "".join(ops) if ops else "*"
        star = rnd.random() < 0.1 or qlen == 0 and ops
        if not ops:
            qlen = rnd.randint(1, 200)
        seq = "*" if star else "".join(rnd.choice("ACGTN") for _ in range(qlen))
        qual = "*" if star or rnd.random() < 0.2 else "".join(chr(rnd.randint(33, 73)) for _ in range(qlen))
        flag = rnd.choice([0, 16, 4, 256, 2048, 83, 163, 77, "0x10", "0x904", "020", "04"])
        rname = rnd.choice(["chr1", "chr1", "chrM", "*", "chr2", "chr"])  # chr2 / chr: not declared -> unmapped
        pos = rnd.choice([0, 1, 5, 99999])
        mapq = rnd.choice([0, 1, 30, 49, 50, 60, 255])
        q = f"read{rnd.randint(0, n // 2)}"
        out.append(f"{q}\t{flag}\t{rname}\t{pos}\t{mapq}\t{cigar}\t=\t{rnd.randint(0, 500)}\t{rnd.randint(-300, 300)}"
                   f"\t{seq}\t{qual}" + ("\tNM:i:1" if rnd.random() < 0.5 else ""))
    eol = "\r\n" if seed % 2 else "\n"
    return (eol.join(out) + (eol if seed % 3 else "")).encode()


def test_sam_cases(golden):
    for c in golden["sam_cases"]:
        want = _ids(c["expect"])
        got = orc.set_from_sam(B(c["buf"]), c["min_len"], c["min_cov"], c["min_mapq"]).sorted_ids()
        assert got == want, c["name"]
        assert sorted(pyo.ids_from_sam(B(c["buf"]), c["min_len"], c["min_cov"], c["min_mapq"])) == want, c["name"]
    for c in golden["sam_errors"]:
        with pytest.raises(orc.OracleError) as e:
            orc.set_from_sam(B(c["buf"]), 0, 0.0, 0)
        assert e.value.code == c["error"], c["name"]
        if "error_line" in c:
            assert e.value.index == c["error_line"], c["name"]
        with pytest.raises(pyo.RefError) as e2:
            pyo.ids_from_sam(B(c["buf"]))
        assert (e2.value.code, e2.value.index) == (e.value.code, e.value.index), c["name"]


@pytest.mark.parametrize("seed", range(6))
def test_sam_two_restatements_agree(seed):
    buf = _random_sam(seed)
    for args in ((0, 0.0, 0), (50, 0.5, 50), (100, 2.0, 30), (10 ** 9, 0.9, 0)):
        assert orc.set_from_sam(buf, *args).sorted_ids() == sorted(pyo.ids_from_sam(buf, *args))


def test_fastq_cases(golden):
    for c in golden["fastq_cases"]:
        ids = [B(i) for i in c["ids"]]
        r = orc.clean_fastq(B(c["buf"]), orc.OSet.from_ids(ids), c["reverse"])
        assert r.written == B(c["written"]), c["name"]
        assert (r.reads_in, r.reads_out) == (c["reads_in"], c["reads_out"]), c["name"]
        assert r.empty_input == c.get("empty_input", False), c["name"]
        w, o, rin, rout = pyo.clean_fastq(B(c["buf"]), set(ids), c["reverse"])
        assert w == B(c["written"]) and (rin, rout) == (c["reads_in"], c["reads_out"]), c["name"]
        assert o == r.other, c["name"]


def test_fastq_errors(golden):
    for c in golden["fastq_errors"]:
        r = orc.clean_fastq(B(c["buf"]), orc.OSet(), False, raise_on_error=False)
        assert (r.error, r.error_record) == (c["error"], c["error_record"]), c["name"]
        with pytest.raises(pyo.RefError) as e:
            pyo.clean_fastq(B(c["buf"]), set())
        assert (e.value.code, e.value.index) == (c["error"], c["error_record"]), c["name"]


def test_diff_cases(golden):
    for c in golden["diff_cases"]:
        pairs = [(B(a), B(b)) for a, b in c["pairs"]]
        rin, rout, d, ids = orc.diff(pairs)
        assert (rin, rout, d) == (c["reads_in"], c["reads_out"], c["difference"]), c["name"]
        assert ids.sorted_ids() == _ids(c["diff_ids"]), c["name"]
        rin, rout, d, ids2 = pyo.diff(pairs)
        assert (rin, rout, d) == (c["reads_in"], c["reads_out"], c["difference"]), c["name"]
        assert sorted(ids2) == _ids(c["diff_ids"])
    rc = golden["report_case"]  # README.md:198-200: reads_removed == reads_in - reads_out
    assert rc["reads_in"] - rc["reads_out"] == rc["reads_removed"]


def test_reads_cases(golden):
    for c in golden["reads_cases"]:
        tx = [B(t) for t in c["taxids"]]
        got = orc.set_from_reads(B(c["buf"]), c["style"], orc.OSet.from_ids(tx)).sorted_ids()
        assert got == _ids(c["expect"]), c["name"]
        assert sorted(pyo.ids_from_reads(B(c["buf"]), c["style"], set(tx))) == _ids(c["expect"]), c["name"]
    for c in golden["reads_errors"]:
        with pytest.raises(orc.OracleError) as e:
            orc.set_from_reads(B(c["buf"]), c["style"], orc.OSet())
        assert e.value.code == c["error"]
        with pytest.raises(pyo.RefError) as e2:
            pyo.ids_from_reads(B(c["buf"]), c["style"], set())
        assert e2.value.code == c["error"]


def test_txt_and_get_id(golden):
    for c in golden["txt_cases"]:
        assert orc.set_from_txt(B(c["buf"])).sorted_ids() == _ids(c["expect"])
        assert sorted(pyo.ids_from_txt(B(c["buf"]))) == _ids(c["expect"])
    for c in golden["get_id_cases"]:
        if "error" in c:
            with pytest.raises(orc.OracleError) as e:
                orc.get_id(B(c["header"]))
            assert e.value.code == c["error"]
            with pytest.raises(pyo.RefError):
                pyo.get_id(B(c["header"]))
        else:
            assert orc.get_id(B(c["header"])) == B(c["expect"])
            assert pyo.get_id(B(c["header"])) == B(c["expect"])


def test_fasta_cases(golden):
    """FASTA input (first byte '>'): both restatements against the hand-derived vectors"""
    for c in golden["fasta_cases"]:
        ids = [B(i) for i in c["ids"]]
        r = orc.clean_fastq(B(c["buf"]), orc.OSet.from_ids(ids), c["reverse"])
        assert (r.written, r.other) == (B(c["written"]), B(c["other"])), c["name"]
        w, o, _, _ = pyo.clean_fastq(B(c["buf"]), set(ids), c["reverse"])
        assert (w, o) == (B(c["written"]), B(c["other"])), c["name"]
    for c in golden["fasta_errors"]:
        r = orc.clean_fastq(B(c["buf"]), orc.OSet(), False, raise_on_error=False)
        assert (r.error, r.error_record) == (c["error"], c["error_record"]), c["name"]
        with pytest.raises(pyo.RefError) as e:
            pyo.clean_fastq(B(c["buf"]), set())
        assert (e.value.code, e.value.index) == (c["error"], c["error_record"]), c["name"]
