"""ctypes wrapper around oracle/scrubby_oracle.c -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  PARITY UNPINNED (see scrubby_oracle.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "scrubby_oracle.c")
    stale = not os.path.exists(_SO) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_SO)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


class Counts(C.Structure):
    _fields_ = [
        ("reads_in", C.c_uint64),
        ("reads_out", C.c_uint64),
        ("difference", C.c_uint64),
        ("error_record", C.c_uint64),
        ("crlf", C.c_uint32),
        ("empty_input", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, u8p, sz, u64 = C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint64
        L.orc_set_new.restype = vp
        L.orc_set_free.argtypes = [vp]
        L.orc_set_insert.argtypes = [vp, u8p, sz]
        L.orc_set_contains.argtypes = [vp, u8p, sz]
        L.orc_set_len.argtypes = [vp]
        L.orc_set_len.restype = u64
        L.orc_set_dump_sorted.argtypes = [vp, C.POINTER(vp), C.POINTER(sz)]
        L.orc_free.argtypes = [vp]
        L.orc_get_id.argtypes = [u8p, sz, C.POINTER(sz), C.POINTER(sz)]
        L.orc_set_from_paf.argtypes = [vp, sz, u64, C.c_double, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]
        L.orc_set_from_sam.argtypes = [vp, sz, u64, C.c_double, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]
        L.orc_set_from_bam.argtypes = [vp, sz, u64, C.c_double, C.c_uint8, C.POINTER(vp), C.POINTER(u64)]
        L.orc_set_from_txt.argtypes = [vp, sz, C.POINTER(vp), C.POINTER(u64)]
        L.orc_taxids_from_report.argtypes = [
            vp, sz, C.POINTER(C.c_char_p), sz, C.POINTER(C.c_char_p), sz, C.POINTER(vp), C.POINTER(u64)]
        L.orc_set_from_reads.argtypes = [vp, sz, C.c_int, vp, C.POINTER(vp), C.POINTER(u64)]
        L.orc_clean_fastq.argtypes = [vp, sz, vp, C.c_int, vp, C.POINTER(sz), vp, C.POINTER(sz), C.POINTER(Counts)]
        L.orc_diff.argtypes = [vp, sz, vp, sz, C.POINTER(Counts), vp]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, code: int, index: int = 0):
        super().__init__(f"oracle error {code} at record/line {index}")
        self.code = code
        self.index = index


def _ptr(buf):
    """address + keep-alive object for bytes / bytearray / numpy uint8 arrays"""
    if isinstance(buf, (bytes, bytearray)):
        arr = (C.c_char * len(buf)).from_buffer_copy(buf) if isinstance(buf, bytes) else (C.c_char * len(buf)).from_buffer(buf)
        return C.cast(arr, C.c_void_p), len(buf), arr
    import numpy as np

    a = np.ascontiguousarray(buf, dtype=np.uint8)
    return C.c_void_p(a.ctypes.data), a.size, a


class OSet:
    """exact string set (HashSet<String>)"""

    def __init__(self, handle=None):
        self.h = C.c_void_p(handle) if handle is not None else C.c_void_p(lib().orc_set_new())

    def __del__(self):
        try:
            if getattr(self, "h", None):
                lib().orc_set_free(self.h)
                self.h = None
        except Exception:  # interpreter shutdown
            pass

    @classmethod
    def from_ids(cls, ids):
        s = cls()
        for i in ids:
            b = i.encode() if isinstance(i, str) else bytes(i)
            lib().orc_set_insert(s.h, b, len(b))
        return s

    def __len__(self):
        return int(lib().orc_set_len(self.h))

    def __contains__(self, key):
        b = key.encode() if isinstance(key, str) else bytes(key)
        return bool(lib().orc_set_contains(self.h, b, len(b)))

    def sorted_ids(self) -> list[bytes]:
        out, n = C.c_void_p(), C.c_size_t()
        lib().orc_set_dump_sorted(self.h, C.byref(out), C.byref(n))
        raw = C.string_at(out, n.value)
        lib().orc_free(out)
        return raw.split(b"\n")[:-1] if raw else []


def get_id(header: bytes) -> bytes:
    off, n = C.c_size_t(), C.c_size_t()
    rc = lib().orc_get_id(header, len(header), C.byref(off), C.byref(n))
    if rc:
        raise OracleError(rc)
    return header[off.value: off.value + n.value]


def set_from_paf(buf, min_len=0, min_cov=0.0, min_mapq=0) -> OSet:
    p, n, keep = _ptr(buf)
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_set_from_paf(p, n, min_len, min_cov, min_mapq, C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


def set_from_sam(buf, min_len=0, min_cov=0.0, min_mapq=0) -> OSet:
    """alignment.rs:117-146 restated for text SAM"""
    p, n, keep = _ptr(buf)
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_set_from_sam(p, n, min_len, min_cov, min_mapq, C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


def set_from_bam(buf, min_len=0, min_cov=0.0, min_mapq=0) -> OSet:
    """alignment.rs:117-146 for binary BAM records; `buf` is the BGZF-decompressed stream"""
    p, n, keep = _ptr(buf)
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_set_from_bam(p, n, min_len, min_cov, min_mapq, C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


def set_from_txt(buf) -> OSet:
    p, n, keep = _ptr(buf)
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_set_from_txt(p, n, C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


def taxids_from_report(buf, taxa, taxa_direct) -> OSet:
    p, n, keep = _ptr(buf)
    ta = (C.c_char_p * max(1, len(taxa)))(*[t.encode() for t in taxa])
    td = (C.c_char_p * max(1, len(taxa_direct)))(*[t.encode() for t in taxa_direct])
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_taxids_from_report(p, n, ta, len(taxa), td, len(taxa_direct), C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


def set_from_reads(buf, style: int, taxids: OSet) -> OSet:
    p, n, keep = _ptr(buf)
    out, err = C.c_void_p(), C.c_uint64()
    rc = lib().orc_set_from_reads(p, n, style, taxids.h, C.byref(out), C.byref(err))
    if rc:
        raise OracleError(rc, err.value)
    return OSet(out.value)


@dataclass
class CleanResult:
    written: bytes
    other: bytes
    reads_in: int
    reads_out: int
    crlf: bool
    empty_input: bool
    error: int = 0
    error_record: int = 0


def clean_fastq(buf, ids: OSet, reverse: bool = False, raise_on_error: bool = True,
                want_bytes: bool = True) -> CleanResult:
    """cleaner.rs:731-760.  `written` is what the reference writes; `other` the complement."""
    import numpy as np

    p, n, keep = _ptr(buf)
    o1 = np.empty(2 * n + 16, dtype=np.uint8)
    o2 = np.empty(2 * n + 16, dtype=np.uint8)
    n1, n2, c = C.c_size_t(), C.c_size_t(), Counts()
    rc = lib().orc_clean_fastq(p, n, ids.h, int(reverse), o1.ctypes.data, C.byref(n1),
                               o2.ctypes.data, C.byref(n2), C.byref(c))
    if rc and raise_on_error:
        raise OracleError(rc, c.error_record)
    return CleanResult(
        o1[: n1.value].tobytes() if want_bytes else o1[: n1.value],
        o2[: n2.value].tobytes() if want_bytes else o2[: n2.value],
        c.reads_in, c.reads_out, bool(c.crlf), bool(c.empty_input), rc, c.error_record)


def diff(pairs, raise_on_error: bool = True):
    """utils.rs:250-285 over [(input_bytes, output_bytes), ...] -> (reads_in, reads_out, difference, OSet)"""
    c = Counts()
    ids = OSet()
    for fin, fout in pairs:
        p1, n1, k1 = _ptr(fin)
        p2, n2, k2 = _ptr(fout)
        rc = lib().orc_diff(p1, n1, p2, n2, C.byref(c), ids.h)
        if rc:
            if raise_on_error:
                raise OracleError(rc, c.error_record)
            return rc, c.error_record
    return c.reads_in, c.reads_out, c.difference, ids
