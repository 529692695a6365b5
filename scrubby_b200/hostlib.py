"""ctypes access to the C++ host logic (scrubby_b200/host -> lib/libscrubby_host.so).

Only the pieces that run on the host by design are exposed here: the Kraken-report taxon
state machine (classifier.rs:124-252; SURVEY F10 "stays on host") and the serde_json float
formatter used by the report writer.  The CLI binary is lib/scrubby.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from . import _lib

HERE = os.path.dirname(os.path.abspath(__file__))
HOST_DIR = os.path.join(HERE, "host")
HOST_SO = os.path.join(HERE, "lib", "libscrubby_host.so")
CLI = os.path.join(HERE, "lib", "scrubby")

_h = None


def build() -> str:
    _lib.load()  # libscrubby_gpu.so first: the host library links against it
    srcs = [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR)]
    stale = not os.path.exists(HOST_SO) or not os.path.exists(CLI) or \
        max(os.path.getmtime(s) for s in srcs) > min(os.path.getmtime(HOST_SO), os.path.getmtime(CLI))
    if stale:
        subprocess.check_call(["make", "-C", HOST_DIR, "-s"])
    return HOST_SO


def load():
    global _h
    if _h is None:
        _lib.load()
        H = C.CDLL(build())
        H.scrubby_host_taxids_from_report.argtypes = [
            C.c_void_p, C.c_size_t, C.POINTER(C.c_char_p), C.c_size_t, C.POINTER(C.c_char_p), C.c_size_t,
            C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint64)]
        H.scrubby_host_free.argtypes = [C.c_void_p]
        H.scrubby_host_free.restype = None
        H.scrubby_host_format_f64.argtypes = [C.c_double, C.c_char_p, C.c_size_t]
        H.scrubby_host_encode_string.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t]
        H.scrubby_host_read_file.argtypes = [C.c_char_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        _h = H
    return _h


class HostError(RuntimeError):
    def __init__(self, kind: int, line: int):
        super().__init__(f"ScrubbyError kind {kind} at report line {line}")
        self.kind, self.line = kind, line


# ScrubbyError::Kind values of scrubby_host.hpp that the report parser can raise
KIND_IO, KIND_KR_PARENT, KIND_KR_READS, KIND_KR_DIRECT, KIND_WOULD_PANIC = 0, 19, 20, 21, 22


def get_taxids_from_report(report: bytes, taxa, taxa_direct) -> list[bytes]:
    """classifier.rs:124-252 on the report's bytes -> sorted taxid strings"""
    H = load()
    ta = (C.c_char_p * max(1, len(taxa)))(*[t.encode() for t in taxa])
    td = (C.c_char_p * max(1, len(taxa_direct)))(*[t.encode() for t in taxa_direct])
    buf = C.create_string_buffer(report, len(report)) if report else C.create_string_buffer(1)
    out, n, err = C.c_void_p(), C.c_size_t(), C.c_uint64()
    rc = H.scrubby_host_taxids_from_report(C.cast(buf, C.c_void_p), len(report), ta, len(taxa), td,
                                           len(taxa_direct), C.byref(out), C.byref(n), C.byref(err))
    if rc:
        raise HostError(rc - 100, err.value)
    raw = C.string_at(out, n.value)
    H.scrubby_host_free(out)
    return raw.split(b"\n")[:-1] if raw else []


def format_f64(v: float) -> str:
    buf = C.create_string_buffer(64)
    n = load().scrubby_host_format_f64(v, buf, 64)
    return buf.raw[:n].decode()


def read_file(path: str) -> bytes:
    """the host stage in front of every parser: gz / BGZF sniffing and inflate (niffler::get_reader's role)"""
    out, n = C.c_void_p(), C.c_size_t()
    rc = load().scrubby_host_read_file(path.encode(), C.byref(out), C.byref(n))
    if rc:
        raise HostError(rc - 100, 0)
    raw = C.string_at(out, n.value)
    load().scrubby_host_free(out)
    return raw


def encode_string(which: int, raw: bytes) -> bytes:
    """which = 0: serde_json string escaping (report JSON); 1: csv field with a tab delimiter (read-id TSV)"""
    buf = C.create_string_buffer(6 * len(raw) + 8)
    n = load().scrubby_host_encode_string(which, raw, len(raw), buf, len(buf))
    assert n >= 0
    return buf.raw[:n]
