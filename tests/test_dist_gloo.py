"""Host logic of the multi-GPU driver (scrubby_b200/dist.py) on CPU: gloo backend, world_size 2 and 3.

The compute calls are served by a CPU stand-in built on the oracle (tests may use the oracle as the
checker); what is under test is the sharding plan, the newline-count all_gather that fixes each shard's
line phase, the collective halo retry, the counter reduction and the output offsets.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from scrubby_b200 import synth  # noqa: E402
from scrubby_b200.dist import Shard, clean_fastq_sharded, plan_shards  # noqa: E402


class HaloError(Exception):
    status = 21  # SGPU_ERR_HALO


class OracleOps:
    """CPU stand-in for GpuOps (TEST ONLY)"""

    def upload(self, b):
        return bytes(b)

    def count_newlines(self, buf, n):
        return buf[:n].count(b"\n")

    def first_line_crlf(self, buf, n):
        p = buf[:n].find(b"\n")
        return p > 0 and buf[p - 1 : p] == b"\r"

    def clean_shard(self, ids, buf, sh: Shard, newlines_before, crlf, reverse, want_other):
        from oracle import oracle as orc

        # records that START inside [0, own_len): a record starts after a newline whose global index is 3 mod 4
        # (a record that starts exactly at a cut belongs to the shard BEFORE the cut: the next one cannot know that its
        # byte 0 follows a newline)
        pos = 0
        if not sh.is_first:
            for _ in range((3 - newlines_before % 4) % 4 + 1):
                p = buf.find(b"\n", pos)
                if p < 0:
                    return b"", b"", 0, 0
                pos = p + 1
        start = pos
        end = start
        while end < sh.own_len or (end == sh.own_len and not sh.is_last and end < len(buf)):
            e = end
            for _ in range(4):
                p = buf.find(b"\n", e)
                if p < 0:
                    if sh.is_last:
                        e = len(buf)
                        break
                    raise HaloError()
                e = p + 1
            end = e
        if start >= end:
            return b"", b"", 0, 0
        r = orc.clean_fastq(buf[start:end], ids, reverse)
        return r.written, (r.other if want_other else b""), r.reads_in, r.reads_out

    # ---- sharded diff (dist.diff_sharded): python sets stand in for device id sets
    def new_set(self):
        return set()

    def set_ids(self, s):
        return sorted(s)

    def set_from_ids(self, id_list):
        return set(id_list)

    def ids_shard(self, probe, buf, sh: Shard, newlines_before, into):
        from oracle import oracle as orc

        pos = 0
        if not sh.is_first:
            for _ in range((3 - newlines_before % 4) % 4 + 1):
                p = buf.find(b"\n", pos)
                if p < 0:
                    return 0, 0
                pos = p + 1
        rec = picked = 0
        while pos < len(buf) and (pos < sh.own_len or (pos == sh.own_len and not sh.is_last)):
            lines = []
            e = pos
            for _ in range(4):
                p = buf.find(b"\n", e)
                if p < 0:
                    if sh.is_last:
                        p = len(buf)
                    else:
                        raise HaloError()
                lines.append(buf[e:p])
                e = p + 1
            if not lines[0]:
                break  # trailing blank line at EOF
            rid = orc.get_id(lines[0][1:].rstrip(b"\r"))
            rec += 1
            if probe is None or rid not in probe:
                picked += 1
                into.add(rid)
            pos = e
        return rec, picked


class OracleOpsUnite(OracleOps):
    """the same stand-in with GpuOps.unite_sets' exchange (key lists as "id\\n" bytes, all-gathered as padded uint8
    tensors, one set build from the concatenation): drives diff_sharded's device-path control flow on gloo"""

    def unite_sets(self, sets, dist):
        from scrubby_b200.dist import _all_gather_ints

        flat = b"".join(k + b"\n" for s in sets for k in sorted(s))
        world = dist.get_world_size() if dist is not None else 1
        if world > 1:
            sizes = [r[0] for r in _all_gather_ints(dist, [len(flat)], world)]
            pad = torch.zeros(max(max(sizes), 16), dtype=torch.uint8)
            if flat:
                pad[: len(flat)] = torch.frombuffer(bytearray(flat), dtype=torch.uint8)
            gathered = [torch.empty_like(pad) for _ in range(world)]
            dist.all_gather(gathered, pad)
            flat = b"".join(bytes(g[:n].numpy()) for g, n in zip(gathered, sizes))
        return set(flat.split(b"\n")[:-1]) if flat else set()


def _diff_worker(rank, world, port, pairs, halo, q, unite=False):
    from scrubby_b200.dist import diff_sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        r = diff_sharded(OracleOpsUnite() if unite else OracleOps(), pairs, dist, halo=halo)
        q.put((rank, r.reads_in, r.reads_out, r.difference, r.diff_ids))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fq, ids_txt, reverse, halo, q):
    from oracle import oracle as orc

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ids = orc.set_from_txt(ids_txt)
        r = clean_fastq_sharded(OracleOps(), ids, fq, dist, reverse=reverse, want_other=True, halo=halo)
        q.put((rank, r.written, r.other, r.offset_written, r.offset_other, r.total_written, r.total_other,
               r.reads_in, r.reads_out))
    finally:
        dist.destroy_process_group()


def _run(world, fq, ids_txt, reverse=False, halo=4096):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, fq, ids_txt, reverse, halo, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


def test_plan_shards_covers_file_exactly():
    for n, w in [(0, 2), (5, 4), (100, 3), (1000, 8), (33, 2), (1 << 20, 8)]:
        sh = plan_shards(n, w, halo=10)
        assert len(sh) == w and sum(s.own_len for s in sh) == n
        assert sum(s.is_last for s in sh) == 1 and sh[0].is_first
        pos = 0
        for s in sh:
            if s.own_len:
                assert s.start == pos and s.start % 16 == 0
                assert s.buf_len >= s.own_len and s.start + s.buf_len <= n
            pos += s.own_len


def test_plan_shards_buffer_at_eof_is_the_last_shard():
    """ADVICE r01 (fastq_records.cuh:53): a buffer that ends at EOF must belong to the shard that applies the end-of-file
    rules -- no non-last shard's halo reaches EOF, the shard whose halo would owns the rest of the file"""
    for n, w, halo in [(1000, 4, 10), (1000, 4, 300), (1000, 4, 2000), (4096, 8, 512), (33, 2, 64), (1 << 20, 8, 1 << 17)]:
        sh = plan_shards(n, w, halo=halo)
        assert sum(s.own_len for s in sh) == n and sum(s.is_last for s in sh) == 1
        for s in sh:
            if s.own_len and not s.is_last:
                assert s.start + s.buf_len < n
            if s.is_last:
                assert s.start + s.buf_len == n and s.buf_len == s.own_len


@pytest.mark.parametrize("tail", ["no_final_newline", "blank_lines_at_cut", "blank_lines"])
def test_sharded_clean_tail_straddles_the_last_cut(tail):
    """the two inputs of ADVICE r01's HALO finding: a last record without a trailing newline that straddles the last
    cut, and trailing blank lines that start exactly at a cut -- accepted by the unsharded path, so by the sharded one"""
    from oracle import oracle as orc

    recs = [b"@t.%d x\n" % i + b"ACGT" * 10 + b"\n+\n" + b"IIII" * 10 + b"\n" for i in range(60)]
    fq = b"".join(recs)
    cases = ((2, 64), (3, 16))
    if tail == "no_final_newline":  # ... on a last record that starts before the last cut and runs to EOF
        fq = b"".join(b"@t%d\nAC\n+\nII\n" % i for i in range(5)) + b"@t.8 x\n" + b"A" * 200 + b"\n+\n" + b"I" * 200
        cases = ((2, 64), (3, 16))
    elif tail == "blank_lines":
        fq += b"\n\n"
    else:  # 96 bytes of records + three blank lines on three ranks: cuts at 48 and 96, the last range is the blank tail
        fq = b"".join(b"@t%d\nAC\n+\nII\n" % i for i in range(8)) + b"\n\n"
        assert len(fq) == 98 and plan_shards(98, 3, 1)[2].start == 96
        cases = ((3, 16), (3, 2))
    ids_txt = b"".join(b"t.%d\nt%d\n" % (i, i) for i in range(0, 60, 2))
    whole = orc.clean_fastq(fq, orc.set_from_txt(ids_txt))
    assert whole.reads_in in (6, 8, 60) and 0 < whole.reads_out < whole.reads_in
    for world, halo in cases:
        res = _run(world, fq, ids_txt, halo=halo)
        assert b"".join(r[1] for r in res) == whole.written
        assert res[0][7] == whole.reads_in and res[0][8] == whole.reads_out


@pytest.mark.parametrize("world,reverse", [(2, False), (2, True), (3, False)])
def test_sharded_clean_matches_unsharded(world, reverse):
    from oracle import oracle as orc

    n = 3000
    fq = synth.gen_fastq(n, 1, start=7).numpy().tobytes()
    ids_txt = synth.gen_txt_ids(n + 10).numpy().tobytes()
    whole = orc.clean_fastq(fq, orc.set_from_txt(ids_txt), reverse)
    res = _run(world, fq, ids_txt, reverse)
    assert b"".join(r[1] for r in res) == whole.written
    assert b"".join(r[2] for r in res) == whole.other
    off_w = off_o = 0
    for r in res:
        assert (r[3], r[4]) == (off_w, off_o)  # write offsets in the concatenated output
        off_w += len(r[1])
        off_o += len(r[2])
        assert (r[5], r[6]) == (len(whole.written), len(whole.other))
        assert (r[7], r[8]) == (whole.reads_in, whole.reads_out)  # counters reduced over ranks


def test_halo_retry_is_collective():
    """records longer than the first halo: every rank retries with a larger one and the result is unchanged"""
    from oracle import oracle as orc

    recs = []
    for i in range(40):
        L = 3000 + 37 * i
        recs.append(b"@long.%d x\n" % i + b"A" * L + b"\n+\n" + b"I" * L + b"\n")
    fq = b"".join(recs)
    ids_txt = b"".join(b"long.%d\n" % i for i in range(0, 40, 3))
    whole = orc.clean_fastq(fq, orc.set_from_txt(ids_txt))
    res = _run(2, fq, ids_txt, halo=64)
    assert b"".join(r[1] for r in res) == whole.written
    assert res[0][7] == whole.reads_in and res[0][8] == whole.reads_out


@pytest.mark.parametrize("world,unite", [(2, False), (3, False), (2, True), (3, True)])
def test_sharded_diff_matches_unsharded(world, unite):
    """dist.diff_sharded: per-rank byte ranges of the output and input files, replicated output-id set, summed
    counters and united id lists equal ReadDifference::get_difference on the whole files (utils.rs:250-285)"""
    from oracle import oracle as orc

    n = 3000
    ids = orc.set_from_txt(synth.gen_txt_ids(n).numpy().tobytes())
    pairs = []
    for mate in (1, 2):
        fq = synth.gen_fastq(n, mate).numpy().tobytes()
        pairs.append((fq, orc.clean_fastq(fq, ids).written))
    want = orc.diff(pairs)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_diff_worker, args=(r, world, port, pairs, 4096, q, unite)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rin, rout, d, dids in res:
        assert (rin, rout, d) == want[:3]
        assert dids == want[3].sorted_ids()


def _error_worker(rank, world, port, fq, q):
    from oracle import oracle as orc

    from scrubby_b200.dist import ShardError

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        try:
            clean_fastq_sharded(OracleOps(), orc.OSet(), fq, dist, halo=4096)
            q.put((rank, "ok", 0))
        except ShardError as e:
            q.put((rank, "shard", e.rank))
        except orc.OracleError as e:
            q.put((rank, "own", e.code))
    finally:
        dist.destroy_process_group()


def test_error_in_one_shard_ends_every_rank():
    """a corrupt record in the last third of the file: the owning rank raises its own error, the others a ShardError
    naming it -- nobody is left waiting in a collective"""
    n = 3000
    fq = bytearray(synth.gen_fastq(n, 1).numpy().tobytes())
    pos = fq.rfind(b"\n@syn.", 0, len(fq) - 1000)  # a header near the end: break its '@'
    fq[pos + 1] = ord("X")
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_error_worker, args=(r, world, port, bytes(fq), q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, "shard", 2), (1, "shard", 2), (2, "own", 3)]  # FASTQ_INVALID_START on the owning rank


# ------------------------------------------------------------------------------------------ one-pass protocol (round 2)
class _SpecResult:
    def __init__(self, status=0, path=1, own_newlines=0, lead_newlines=0, crlf=False, written=b"", other=b"", reads_in=0,
                 reads_out=0):
        self.status, self.path, self.own_newlines, self.lead_newlines, self.crlf = status, path, own_newlines, lead_newlines, crlf
        self.written, self.other = written, other
        self.n_written, self.n_other, self.reads_in, self.reads_out = len(written), len(other), reads_in, reads_out


def _spec_call(ops, ids, buf, sh, nb, crlf, reverse):
    """CPU stand-in (TEST ONLY) for sgpu_clean_fastq_shard_dev incl. SGPU_NEWLINES_UNKNOWN: the speculation rule of
    fastq_fused.cu::fused_shard restated -- among the first four newlines the first one followed by "+\\n" ends a
    sequence line; canonical LF input only, anything else answers SGPU_ERR_PHASE_UNKNOWN (24)"""
    own_nl = buf[: sh.own_len].count(b"\n")
    if nb is None:
        if b"\r" in buf or not sh.own_len:
            return _SpecResult(status=24)
        pos, p = [], -1
        for _ in range(4):
            p = buf.find(b"\n", p + 1)
            if p < 0:
                break
            pos.append(p)
        j = next((i for i, p in enumerate(pos) if buf[p + 1: p + 3] == b"+\n"), None)
        if j is None:
            return _SpecResult(status=24)
        k = (2 + j) & 3
        if k >= len(pos) or pos[k] + 1 > sh.own_len or pos[k] + 1 >= len(buf):
            return _SpecResult(status=24)
        lead = k + 1
        try:
            w, o, rin, rout = ops.clean_shard(ids, buf, sh, (4 - lead % 4) % 4, False, reverse, True)
        except Exception:  # a wrong phase makes the records malformed: the single-pass kernel falls back -> 24
            return _SpecResult(status=24)
        return _SpecResult(0, 1, own_nl, lead, False, w, o, rin, rout)
    first_crlf = ops.first_line_crlf(buf, len(buf)) if sh.is_first and crlf is None else bool(crlf)
    if first_crlf or b"\r" in buf:
        # (the oracle's whole-file cleaner normalises line endings; shards of CRLF files are not needed by these tests)
        w, o, rin, rout = ops.clean_shard(ids, buf, sh, nb, first_crlf, reverse, True)
        return _SpecResult(0, 2, own_nl, 0, first_crlf, w, o, rin, rout)
    w, o, rin, rout = ops.clean_shard(ids, buf, sh, nb, False, reverse, True)
    return _SpecResult(0, 1, own_nl, 0, False, w, o, rin, rout)


def _onepass_worker(rank, world, port, files, ids_txt, halo, q):
    from oracle import oracle as orc

    from scrubby_b200.dist import _clean_files_sharded

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ids = orc.set_from_txt(ids_txt)
        ops = OracleOps()
        shards = [plan_shards(len(f), world, halo)[rank] for f in files]
        bufs = [f[sh.start: sh.start + sh.buf_len] for f, sh in zip(files, shards)]
        last = {}

        def call(f, nb, crlf):
            last[f] = _spec_call(ops, ids, bufs[f], shards[f], nb, crlf, False)
            return last[f]

        rs = _clean_files_sharded(call, lambda f: bufs[f][: shards[f].own_len].count(b"\n"),
                                  lambda f: ops.first_line_crlf(bufs[f], len(bufs[f])), shards, dist, "cpu", HaloError)
        q.put((rank, [(last[f].written if f in last and last[f].status == 0 else b"", r.offset_written, r.total_written,
                       r.reads_in, r.reads_out, r.one_pass) for f, r in enumerate(rs)]))
    finally:
        dist.destroy_process_group()


def _run_onepass(world, files, ids_txt, halo=4096):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_onepass_worker, args=(r, world, port, files, ids_txt, halo, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("world", [2, 3])
def test_one_pass_protocol_accepts_canonical_files(world):
    """both mate files in one exchange: every speculation is accepted, offsets and counters equal the unsharded run"""
    from oracle import oracle as orc

    n = 3000
    files = [synth.gen_fastq(n, m, start=3).numpy().tobytes() for m in (1, 2)]
    ids_txt = synth.gen_txt_ids(n + 10).numpy().tobytes()
    res = _run_onepass(world, files, ids_txt)
    for f, fq in enumerate(files):
        whole = orc.clean_fastq(fq, orc.set_from_txt(ids_txt))
        assert b"".join(r[1][f][0] for r in res) == whole.written
        off = 0
        for r in res:
            w, offset, total, rin, rout, one_pass = r[1][f]
            assert one_pass and offset == off and total == len(whole.written)
            assert (rin, rout) == (whole.reads_in, whole.reads_out)
            off += len(w)


def test_one_pass_protocol_refutes_wrong_speculation():
    """one-base reads whose quality line is "+": the first "\\n+\\n" of a shard may be a QUALITY line, the speculated phase
    is then wrong, the exchanged newline counts refute it and the exact protocol produces the right bytes"""
    from oracle import oracle as orc

    fq = None
    for pad in range(64):  # shift the file until the world-2 cut lands on the '+' of a separator before a "+" quality line
        recs = [b"@r%d%s\nA\n+\n%s\n" % (i, b"x" * pad if i == 0 else b"", b"+" if i % 3 else b"I") for i in range(4000)]
        cand = b"".join(recs)
        cut = plan_shards(len(cand), 2, 256)[1].start
        if cand[cut - 1: cut + 4] == b"\n+\n+\n":
            fq = cand
            break
    assert fq is not None
    ids_txt = b"".join(b"r%d\n" % i for i in range(0, 4000, 2))
    whole = orc.clean_fastq(fq, orc.set_from_txt(ids_txt))
    for world in (2, 3):
        res = _run_onepass(world, [fq], ids_txt, halo=256)
        assert b"".join(r[1][0][0] for r in res) == whole.written
        assert all(r[1][0][3:5] == (whole.reads_in, whole.reads_out) for r in res)
        if world == 2:
            assert not res[0][1][0][5], "the cut was built to refute the speculation"
