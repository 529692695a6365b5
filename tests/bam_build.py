"""Minimal BAM writer for the tests (SAMv1 section 4.2): records -> the uncompressed BAM stream, optionally wrapped
in BGZF blocks (concatenated gzip members with the BC extra field + the EOF marker block)."""
import struct
import zlib

OPS = "MIDNSHP=XB"


def cigar_ops(cigar: str):
    out, num = [], ""
    for ch in cigar:
        if ch.isdigit():
            num += ch
        else:
            out.append((int(num) << 4) | OPS.index(ch))
            num = ""
    return out


def record(qname: bytes, flag=0, ref_id=0, pos=100, mapq=60, cigar="", l_seq=None, aux=b"", nul=True, ops=None,
           block_size_delta=0) -> bytes:
    ops = cigar_ops(cigar) if ops is None else ops
    if l_seq is None:
        l_seq = sum(v >> 4 for v in ops if (v & 15) in (0, 1, 4, 7, 8))
    name = qname + (b"\x00" if nul else b"")
    body = struct.pack("<iiBBHHHiiii", ref_id, pos, len(name), mapq, 4680, len(ops), flag, l_seq, -1, -1, 0)
    body += name + b"".join(struct.pack("<I", v) for v in ops) + b"\x11" * ((l_seq + 1) // 2) + b"\x28" * l_seq + aux
    return struct.pack("<I", len(body) + block_size_delta) + body


def stream(records, refs=((b"chr1", 1000000),), text=b"@HD\tVN:1.6\n") -> bytes:
    out = b"BAM\x01" + struct.pack("<I", len(text)) + text + struct.pack("<I", len(refs))
    for name, length in refs:
        out += struct.pack("<I", len(name) + 1) + name + b"\x00" + struct.pack("<I", length)
    return out + b"".join(records)


def bgzf(data: bytes, block: int = 0xFF00) -> bytes:
    out = b""
    for o in list(range(0, len(data), block)) + [None]:
        chunk = b"" if o is None else data[o: o + block]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        bsize = 12 + 6 + len(comp) + 8
        out += b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
        out += comp + struct.pack("<II", zlib.crc32(chunk), len(chunk))
    return out
