"""Second, independently written restatement of the hot path in plain Python.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see scrubby_oracle.h).  Small inputs
only; exists so that the C oracle is checked by something other than itself.
Every function cites the reference file:line it follows (/root/reference/src).
"""
from __future__ import annotations

# error classes share numbering with oracle/scrubby_oracle.h
E_IO, E_START, E_SEP, E_UNEQ, E_END, E_FMT, E_UTF8, E_HEADER = 1, 3, 4, 5, 6, 7, 8, 9
E_PAFINT, E_PANIC, E_KR_READS, E_KR_DIRECT, E_KR_PARENT, E_FASTA = 10, 11, 12, 13, 14, 15


class RefError(Exception):
    def __init__(self, code, index=0):
        super().__init__(f"error {code} at {index}")
        self.code, self.index = code, index


# Unicode White_Space, the set Rust's char::is_whitespace uses.
_WS = {0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F, 0x205F, 0x3000}
_WS |= set(range(0x2000, 0x200B))


def _split_whitespace(s: str) -> list[str]:
    out, cur = [], []
    for ch in s:
        if ord(ch) in _WS:
            if cur:
                out.append("".join(cur))
                cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return out


def _trim(s: str) -> str:
    i, j = 0, len(s)
    while i < j and ord(s[i]) in _WS:
        i += 1
    while j > i and ord(s[j - 1]) in _WS:
        j -= 1
    return s[i:j]


def get_id(header: bytes) -> bytes:
    """utils.rs:91-103"""
    try:
        text = header.decode("utf-8")  # strict: same acceptance as Rust from_utf8
    except UnicodeDecodeError:
        raise RefError(E_UTF8)
    parts = _split_whitespace(text)
    if len(parts) < 1:
        raise RefError(E_HEADER)
    return parts[0].encode("utf-8")


def _lines(buf: bytes):
    """BufRead::lines: split on LF, strip one CR, every line must be UTF-8."""
    pos, no = 0, 0
    while pos < len(buf):
        nl = buf.find(b"\n", pos)
        raw = buf[pos:] if nl < 0 else buf[pos: nl + 1]
        pos += len(raw)
        try:
            text = raw.decode("utf-8")
        except UnicodeDecodeError:
            raise RefError(E_IO, no)
        if text.endswith("\n"):
            text = text[:-1]
            if text.endswith("\r"):
                text = text[:-1]
        yield no, text
        no += 1


def _rust_uint(s: str, bits: int) -> int:
    """<uN as FromStr>: optional '+', ASCII digits only, overflow rejected."""
    if s == "":
        raise ValueError
    if s[0] in "+-":
        if len(s) == 1 or s[0] == "-":
            raise ValueError
        s = s[1:]
    v = 0
    for ch in s:
        if not ("0" <= ch <= "9"):
            raise ValueError
        v = v * 10 + (ord(ch) - 48)
        if v >= (1 << bits):
            raise ValueError
    return v


def ids_from_paf(buf: bytes, min_len=0, min_cov=0.0, min_mapq=0) -> set[bytes]:
    """alignment.rs:84-114 with PafRecord::from_str :244-263 and :265-275"""
    ids = set()
    kinds = {1: 64, 2: 64, 3: 64, 6: 64, 7: 64, 8: 64, 9: 64, 10: 64, 11: 8}
    for no, line in _lines(buf):
        f = line.split("\t")
        v = {}
        for c in range(12):
            if c >= len(f):
                raise RefError(E_PANIC, no)
            if c in kinds:
                try:
                    v[c] = _rust_uint(f[c], kinds[c])
                except ValueError:
                    raise RefError(E_PAFINT, no)
        alen = (v[3] - v[2]) % (1 << 64)
        cov = 0.0 if v[1] == 0 else float(alen) / float(v[1])
        if (alen >= min_len or cov >= min_cov) and v[11] >= min_mapq:
            ids.add(f[0].encode("utf-8"))
    return ids


E_SAM = 22


def ids_from_sam(buf: bytes, min_len=0, min_cov=0.0, min_mapq=0) -> set[bytes]:
    """alignment.rs:117-146 + :154-211 for text SAM (second, independent restatement of the rules listed in
    scrubby_oracle.c; regular expressions instead of scanners)"""
    import re

    ids = set()
    raw = buf.split(b"\n")
    if raw and raw[-1] == b"":
        raw.pop()
        terminated = [True] * len(raw)
    else:
        terminated = [True] * (len(raw) - 1) + [False]
    # the reference names the header declares (SN: of every @SQ line): htslib looks RNAME up among them
    refs = set()
    for line, term in zip(raw, terminated):
        if term and line.endswith(b"\r"):
            line = line[:-1]
        if line.startswith(b"@SQ\t"):
            sn = [x[3:] for x in line.split(b"\t")[1:] if x.startswith(b"SN:")]
            if sn:
                refs.add(sn[0])
    have_sq = any((l[:-1] if (t and l.endswith(b"\r")) else l).startswith(b"@SQ\t") and
                  any(x.startswith(b"SN:") for x in (l[:-1] if (t and l.endswith(b"\r")) else l).split(b"\t")[1:])
                  for l, t in zip(raw, terminated))
    for no, (line, term) in enumerate(zip(raw, terminated)):
        if term and line.endswith(b"\r"):
            line = line[:-1]
        if line.startswith(b"@"):
            continue
        f = line.split(b"\t")
        if len(f) < 11:
            raise RefError(E_SAM, no)
        q, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = f[:11]
        m = re.fullmatch(rb"0[xX]([0-9a-fA-F]+)|0([0-7]+)|([1-9][0-9]*|0)", flag)
        sint = lambda x: re.fullmatch(rb"[+-]?[0-9]{1,18}", x)
        ops = re.findall(rb"([0-9]+)([MIDNSHP=XB])", cigar)
        cigar_ok = cigar == b"*" or (ops and b"".join(a + b for a, b in ops) == cigar and
                                     all(int(a) < (1 << 28) for a, _ in ops))
        if not (q and m and rname and (rname == b"*" or have_sq) and sint(pos) and
                re.fullmatch(rb"[0-9]{1,18}", mapq) and int(mapq) <= 255 and
                cigar_ok and rnext and sint(pnext) and sint(tlen) and seq and qual):
            raise RefError(E_SAM, no)
        fl = int(m.group(1), 16) if m.group(1) else int(m.group(2), 8) if m.group(2) else int(m.group(3))
        if fl > 65535:
            raise RefError(E_SAM, no)
        if cigar == b"*":
            ops = []
        qlen = 0 if seq == b"*" else len(seq)
        cq = sum(int(a) for a, o in ops if o in b"MIS=X") & 0xFFFFFFFF
        qalen = sum(int(a) for a, o in ops if o in b"MI") & 0xFFFFFFFF
        if (ops and seq != b"*" and cq != qlen) or (qual != b"*" and len(qual) != qlen):
            raise RefError(E_SAM, no)
        try:
            q.decode("utf-8")
        except UnicodeDecodeError:
            raise RefError(E_UTF8, no)
        if (fl & 4) or rname == b"*" or rname not in refs or int(pos) < 1:
            continue
        cov = 0.0 if qlen == 0 else float(qalen) / float(qlen)
        if (qalen >= min_len or cov >= min_cov) and int(mapq) >= min_mapq:
            ids.add(q)
    return ids


E_BAM = 23


def ids_from_bam(buf: bytes, min_len=0, min_cov=0.0, min_mapq=0) -> set[bytes]:
    """alignment.rs:117-146 + BamRecord::from :180-197 over the BGZF-decompressed BAM stream (second, independent
    restatement of the rules listed in scrubby_oracle.c: struct.unpack and slices instead of pointer walks)"""
    import struct

    def need(cond, rec=0):
        if not cond:
            raise RefError(E_BAM, rec)

    need(len(buf) >= 12 and buf[:4] == b"BAM\x01")
    (l_text,) = struct.unpack_from("<I", buf, 4)
    pos = 8 + l_text
    need(pos + 4 <= len(buf))
    (n_ref,) = struct.unpack_from("<I", buf, pos)
    pos += 4
    for _ in range(n_ref):
        need(pos + 4 <= len(buf))
        (l_name,) = struct.unpack_from("<I", buf, pos)
        pos += 4 + l_name + 4
        need(pos <= len(buf))
    ids, rec = set(), 0
    while pos < len(buf):
        need(pos + 4 <= len(buf), rec)
        (bs,) = struct.unpack_from("<I", buf, pos)
        need(bs >= 32 and pos + 4 + bs <= len(buf), rec)
        blk = buf[pos + 4: pos + 4 + bs]
        ref_id, rpos, l_name, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", blk, 0)
        need(l_name >= 1 and l_seq >= 0 and 32 + l_name + 4 * n_cig + (l_seq + 1) // 2 + l_seq <= bs, rec)
        pos += 4 + bs
        rec += 1
        if flag & 4:
            continue
        name = blk[32: 32 + l_name]
        qname = name[:-1] if name.endswith(b"\x00") else name
        try:
            qname.decode("utf-8")
        except UnicodeDecodeError:
            raise RefError(E_UTF8, rec - 1)
        c0 = 32 + l_name
        ops = list(struct.unpack_from("<%dI" % n_cig, blk, c0))
        if ops and ref_id >= 0 and rpos >= 0 and ops[0] & 15 == 4 and ops[0] >> 4 == l_seq:
            cg = _bam_cg(blk[c0 + 4 * n_cig + (l_seq + 1) // 2 + l_seq:])
            if cg is not None and n_cig <= len(cg) < (1 << 29):
                ops = cg
        qalen = sum(v >> 4 for v in ops if v & 15 in (0, 1)) & 0xFFFFFFFF
        cov = 0.0 if l_seq == 0 else float(qalen) / float(l_seq)
        if (qalen >= min_len or cov >= min_cov) and mapq >= min_mapq:
            ids.add(bytes(qname))
    return ids


def _bam_cg(aux: bytes):
    """the CG:B,I (or B,i) array of a record's auxiliary fields, or None (htslib bam_aux_get + bam_tag2cigar)"""
    import struct

    fixed = {"A": 1, "c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4, "d": 8}
    p = 0
    while len(aux) - p >= 3:
        tag, ty = aux[p: p + 2], chr(aux[p + 2])
        p += 3
        if ty in fixed:
            size = fixed[ty]
        elif ty in "ZH":
            z = aux.find(b"\x00", p)
            if z < 0:
                return None
            size = z - p + 1
        elif ty == "B":
            if len(aux) - p < 5:
                return None
            sub, cnt = chr(aux[p]), struct.unpack_from("<I", aux, p + 1)[0]
            es = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}.get(sub)
            if es is None or cnt * es > len(aux) - p - 5:
                return None
            if tag == b"CG":
                return list(struct.unpack_from("<%dI" % cnt, aux, p + 5)) if sub in "Ii" else None
            size = 5 + cnt * es
        else:
            return None
        if tag == b"CG" or len(aux) - p < size:
            return None
        p += size
    return None


def ids_from_txt(buf: bytes) -> set[bytes]:
    """alignment.rs:60-82"""
    return {line.encode("utf-8") for _, line in _lines(buf)}


_LEVELS = ["None", "Unclassified", "NoRank", "Root", "Domain", "Kingdom", "Phylum", "Class",
           "Order", "Family", "Genus", "Species", "Unspecified"]


def _level(code: str) -> int:
    """classifier.rs:345-373"""
    table = [("U", None, 1), ("no rank", None, 2), ("R", None, 3), ("D", "superkingdom", 4),
             ("K", "kingdom", 5), ("P", "phylum", 6), ("C", "class", 7), ("O", "order", 8),
             ("F", "family", 9), ("G", "genus", 10), ("S", "species", 11)]
    for a, b, lv in table:
        if code.startswith(a) or (b is not None and code.startswith(b)):
            return lv
    return 12


def taxids_from_report(buf: bytes, taxa, taxa_direct) -> set[bytes]:
    """classifier.rs:124-252"""
    taxa = [_trim(t) for t in taxa]
    direct = [_trim(t) for t in taxa_direct]
    out = set()
    level_on, parent = 0, ""
    for no, line in _lines(buf):
        f = line.split("\t")
        if len(f) < 2:
            raise RefError(E_PANIC, no)
        try:
            _rust_uint(f[1], 64)
        except ValueError:
            raise RefError(E_KR_READS, no)
        if len(f) < 3:
            raise RefError(E_PANIC, no)
        try:
            n_direct = _rust_uint(f[2], 64)
        except ValueError:
            raise RefError(E_KR_DIRECT, no)
        if len(f) < 6:
            raise RefError(E_PANIC, no)
        code, tid, name = _trim(f[3]), _trim(f[4]), _trim(f[5])
        lv = _level(code)
        if name in direct or tid in direct:
            out.add(tid.encode())
        if lv < 4:
            continue
        if name in taxa or tid in taxa:
            level_on, parent = lv, name
            if n_direct > 0:
                out.add(tid.encode())
        else:
            if level_on == 0:
                continue
            if lv <= level_on and len(code) == 1:
                level_on = 0
            elif n_direct > 0:
                out.add(tid.encode())
                if parent == "":
                    raise RefError(E_KR_PARENT, no)
    return out


def ids_from_reads(buf: bytes, style: int, taxids: set[bytes]) -> set[bytes]:
    """classifier.rs:270-290 (kraken2, style 0) / :308-328 (metabuli, style 1)"""
    need = 5 if style == 0 else 7
    out = set()
    for no, line in _lines(buf):
        f = line.split("\t")
        if len(f) < need:
            raise RefError(E_PANIC, no)
        if _trim(f[2]).encode("utf-8") in taxids:
            out.add(_trim(f[1]).encode("utf-8"))
    return out


def _trim_cr(b: bytes) -> bytes:
    return b[:-1] if b.endswith(b"\r") else b


def fastq_records(buf: bytes):
    """needletail 0.5.1 fastq Reader::next: yields (index, id, seq, qual, crlf_of_first_record).

    Raises RefError with the index of the failing record."""
    if len(buf) < 5:  # utils.rs:359-375 via niffler FileTooShort
        return
    if buf[:1] == b">":
        yield from fasta_records(buf)
        return
    if buf[:1] != b"@":
        raise RefError(E_FMT)
    start, idx, crlf = 0, 0, None
    n = len(buf)
    while True:
        nls, frm = [], start
        while len(nls) < 4:
            p = buf.find(b"\n", frm)
            if p < 0:
                break
            nls.append(p)
            frm = p + 1
        last = False
        if len(nls) == 4:
            end = nls[3]
        elif len(nls) == 3:
            end, last = n, True
        else:
            rest = buf[start:]
            if all(_trim_cr(l) == b"" for l in rest.split(b"\n")):
                return
            raise RefError(E_END, idx)
        if buf[start: start + 1] != b"@":
            raise RefError(E_START, idx)
        if buf[nls[1] + 1: nls[1] + 2] != b"+":
            raise RefError(E_SEP, idx)
        rid = _trim_cr(buf[start + 1: nls[0]])
        seq = _trim_cr(buf[nls[0] + 1: nls[1]])
        qual = _trim_cr(buf[nls[2] + 1: end])
        if len(seq) != len(qual):
            raise RefError(E_UNEQ, idx)
        if crlf is None:
            crlf = nls[0] > start and buf[nls[0] - 1: nls[0]] == b"\r"
        yield idx, rid, seq, qual, crlf
        idx += 1
        if last:
            return
        start = end + 1


def fasta_records(buf: bytes):
    """needletail 0.5.1 fasta Reader::next (second restatement: split on "\\n>" instead of walking newlines): yields
    (index, id, raw_seq, None, crlf_known_so_far).  raw_seq keeps the inner line breaks of a multi-line sequence."""
    chunks = buf.split(b"\n>")
    crlf = None
    for idx, ch in enumerate(chunks):
        body = ch[1:] if idx == 0 else ch           # without the leading '>'
        last_rec = idx == len(chunks) - 1
        # the newline positions the reference records ("seq_pos"), relative to `body`
        pos = [i for i, c in enumerate(body) if c == 10]
        if last_rec:
            if pos and pos[-1] == len(body) - 1:    # a newline on the input's last byte only counts after another one
                pos = pos if len(pos) > 1 else []
            elif pos:
                pos.append(len(body))               # no trailing newline: the end of the input closes the last line
        else:
            pos.append(len(body))                   # the newline in front of the next '>'
        if not pos:
            raise RefError(E_END, idx)
        rid = _trim_cr(body[: pos[0]])
        raw = _trim_cr(body[pos[0] + 1: pos[-1]]) if len(pos) > 1 else b""
        if crlf is None and 10 in body[: pos[-1]]:
            crlf = body[: pos[-1]].find(b"\n") > 0 and body[body[: pos[-1]].find(b"\n") - 1] == 13
        yield idx, rid, raw, None, bool(crlf)


def clean_fastq(buf: bytes, ids: set[bytes], reverse=False):
    """cleaner.rs:731-760 -> (written, other, reads_in, reads_out).  Raises RefError."""
    w, o = bytearray(), bytearray()
    rin = rout = 0
    for idx, rid, seq, qual, crlf in fastq_records(buf):
        try:
            key = get_id(rid)
        except RefError as e:
            raise RefError(e.code, idx)
        e_ = b"\r\n" if crlf else b"\n"
        rec = b"@" + rid + e_ + seq + e_ + b"+" + e_ + qual + e_ if qual is not None else b">" + rid + e_ + seq + e_
        rin += 1
        hit = key in ids
        if (not reverse and not hit) or (reverse and hit):
            w += rec
            rout += 1
        else:
            o += rec
    return bytes(w), bytes(o), rin, rout


def diff(pairs):
    """utils.rs:250-285 -> (reads_in, reads_out, difference, diff_ids)"""
    diff_ids = set()
    rin = rout = d = 0
    for fin, fout in pairs:
        o_ids = set()
        def key(rid, idx):
            try:
                return get_id(rid)
            except RefError as e:  # the failing record's index within its file
                raise RefError(e.code, idx)

        for idx, rid, *_ in fastq_records(fout):
            o_ids.add(key(rid, idx))
            rout += 1
        for idx, rid, *_ in fastq_records(fin):
            k = key(rid, idx)
            if k not in o_ids:
                diff_ids.add(k)
                d += 1
            rin += 1
    return rin, rout, d, diff_ids
