"""Real NCCL check of the multi-GPU driver (run under torchrun on >= 2 GPUs; tests/test_gpu_multi.py launches it from
pytest and self-skips below two devices):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
        tests/nccl_check.py

Rank 0 builds the id set from Kraken2 lines, the set is replicated by NCCL broadcast, every rank cleans its
byte range of both mate files, counters are allreduced, and the rank-order concatenation must be byte-identical
to the unsharded single-GPU output (which tests/test_gpu_parity.py pins to the oracle).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench import taxids_for_config
from scrubby_b200 import api, synth
from scrubby_b200 import dist as sdist
from scrubby_b200.dist import GpuOps, clean_fastq_sharded, diff_sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = api.Context(local)
ops = GpuOps(ctx)
n = 300_000
taxids = taxids_for_config()
ids = api.IdSet.from_reads(ctx, synth.gen_kraken_reads(n).numpy().tobytes(), 0, taxids) if rank == 0 else None
ids = ops.replicate_set(ids, dist, 0)
ok = True
for mate in (1, 2):
    fq = synth.gen_fastq(n, mate).numpy().tobytes()
    for reverse in (False, True):
        r = clean_fastq_sharded(ops, ids, fq, dist, reverse=reverse, want_other=True)
        parts = [None] * world
        dist.all_gather_object(parts, (r.written, r.other))
        if rank == 0:
            whole = api.clean_fastq(ctx, ids, fq, reverse)
            cat_w = b"".join(p[0] for p in parts)
            cat_o = b"".join(p[1] for p in parts)
            good = cat_w == whole.written and cat_o == whole.other and \
                (r.reads_in, r.reads_out) == (whole.reads_in, whole.reads_out) and r.total_written == len(whole.written)
            print(f"mate {mate} reverse {reverse}: world {world} reads_in {r.reads_in} reads_out {r.reads_out} "
                  f"bytes {len(cat_w)} -> {'identical' if good else 'MISMATCH'}")
            ok = ok and good
# ---- config 5: sharded diff (output-id set all-gathered, counters summed, absent ids united) vs one GPU
pairs = []
for mate in (1, 2):
    fq = synth.gen_fastq(n, mate).numpy().tobytes()
    pairs.append((fq, api.clean_fastq(ctx, ids, fq, False).written))
sd = diff_sharded(ops, pairs, dist)
if rank == 0:
    whole = api.diff(ctx, pairs)
    good = (sd.reads_in, sd.reads_out, sd.difference) == whole[:3] and sd.diff_ids == whole[3].sorted_ids()
    print(f"diff: world {world} reads_in {sd.reads_in} reads_out {sd.reads_out} difference {sd.difference} "
          f"unique ids {len(sd.diff_ids)} -> {'identical' if good else 'MISMATCH'}")
    ok = ok and good
# ---- the device-resident plane bench.py times (round 2): the id list all-gathered as equal byte ranges, the set built
#      on every rank, both mate files in ONE pass with a speculated line phase and one exchange of a few integers;
#      device buffers and host (pinned) buffers; concatenation in rank order == the single-GPU output
txt = synth.gen_txt_ids(n, device=dev)
total = int(txt.numel())
per = sdist.evidence_shard_len(total, world)
d_ev = torch.zeros(per + 16, dtype=torch.uint8, device=dev)
mine = txt[rank * per: min(total, (rank + 1) * per)]
d_ev[: mine.numel()] = mine
ev = sdist.replicate_file_dev(d_ev, per, total, dist)
assert torch.equal(ev, txt), "the all-gathered byte ranges are the file"
ids2 = api.IdSet.from_txt(ctx, ev)
files = [synth.gen_fastq(n, mate, device=dev) for mate in (1, 2)]
whole = [api.clean_fastq(ctx, ids2, f.cpu().numpy().tobytes()) for f in files]
for host in (False, True):
    jobs, keep = [], []
    for f in files:
        sh = sdist.plan_shards(int(f.numel()), world, halo=1 << 16)[rank]
        buf = torch.zeros(sh.buf_len + 16, dtype=torch.uint8, device=dev)
        buf[: sh.buf_len] = f[sh.start: sh.start + sh.buf_len]
        if host:
            hb = torch.empty(sh.buf_len + 16, dtype=torch.uint8, pin_memory=True)
            hb.copy_(buf)
            jobs.append((hb, sh, torch.empty(sh.buf_len + 64, dtype=torch.uint8, pin_memory=True),
                         torch.empty(sh.buf_len + 64, dtype=torch.uint8, pin_memory=True)))
        else:
            jobs.append((buf, sh, torch.empty(sh.buf_len + 64, dtype=torch.uint8, device=dev),
                         torch.empty(sh.buf_len + 64, dtype=torch.uint8, device=dev)))
    rs = (sdist.clean_files_sharded_host if host else sdist.clean_files_sharded_dev)(api, ctx, ids2, jobs, dist)
    for k, (r, job) in enumerate(zip(rs, jobs)):
        mine_w = bytes(job[2][: r.n_written].cpu().numpy())
        mine_o = bytes(job[3][: r.n_other].cpu().numpy())
        parts = [None] * world
        dist.all_gather_object(parts, (mine_w, mine_o, r.offset_written))
        if rank == 0:
            cat_w, cat_o = b"".join(p[0] for p in parts), b"".join(p[1] for p in parts)
            offs = [p[2] for p in parts]
            good = cat_w == whole[k].written and cat_o == whole[k].other and r.one_pass and \
                (r.reads_in, r.reads_out) == (whole[k].reads_in, whole[k].reads_out) and \
                offs == [sum(len(p[0]) for p in parts[:q]) for q in range(world)]
            print(f"one-pass {'host' if host else 'device'} buffers, mate {k + 1}: world {world} reads_in {r.reads_in} "
                  f"reads_out {r.reads_out} one_pass {r.one_pass} -> {'identical' if good else 'MISMATCH'}")
            ok = ok and good
# ---- sharded set build (ShardedTxtSet): every rank partitions its byte range of the id list by virtual page, the lists
#      are read over NVLink (directly, or pulled first), every rank assembles the same table: == from_txt on the whole file
want_ids = ids2.sorted_ids()
for direct in (True, False):
    for lst, name in ((txt, "inline ids"), (synth.gen_txt_ids_illumina(20000, device=dev), "long ids (falls back)")):
        tot = int(lst.numel())
        per2 = sdist.evidence_shard_len(tot, world)
        sh = sdist.ShardedTxtSet(api, ctx, dist, tot, per2, dev, direct=direct)
        for rep in range(3):  # the symmetric buffers are double-buffered: several builds in a row
            buf, nb, own, starts, last = sdist.evidence_shard_with_halo(lst, rank, world, per2, halo=4096)
            got = sh.build(buf, nb, own, starts, last)
            if name == "inline ids":
                good = got is not None and got.sorted_ids() == want_ids
            else:
                good = got is None
            if rank == 0 and rep == 0:
                print(f"sharded set build ({'direct peer reads' if direct else 'pulled lists'}), {name}: world {world} "
                      f"-> {'identical' if good else 'MISMATCH'}")
            t = torch.tensor([1 if good else 0], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = ok and bool(int(t))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.broadcast(flag, 0)
dist.destroy_process_group()
if rank == 0:
    print("nccl_check", "ok" if ok else "FAILED")
sys.exit(0 if int(flag) else 1)
