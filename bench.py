#!/usr/bin/env python
"""bench.py -- headline benchmark of the depletion hot path (BASELINE.json: reads/s & FASTQ GB/s depleted).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P] [--split]

One "step" = one pass of the hot path over one batch of synthetic input:
    Kraken2 per-read lines -> taxid bitmap test -> exact read-id set  (sgpu_idset_from_reads)
    R1 and R2 FASTQ -> parse -> probe -> compact                      (sgpu_clean_fastq, fused kernel)
Workload at N=1: BASELINE.json configs[1] ("scrubby classifier: synthetic 10M pairs + Kraken2 reads/report,
-T Chordata -D 9606, deplete").  N>1 is weak scaling: every rank owns its own 10M-pair shard (reads shard by
chunk; Kraken2 emits its per-read lines in read order, so the evidence is co-partitioned), report counters
are NCCL-allreduced, the timed region is bracketed by barriers and the max over ranks is taken.

`value`  : reads/s with every input already resident in HBM (CUDA events on the launching stream).
`e2e`    : the same metric through the host-buffer C ABI call (pinned host memory; H2D + D2H inside).
`roofline`: the fused kernel's algorithmic bytes (input + output bytes) / its CUDA-event time.
`cpu_baseline`: the CPU oracle (a port structured like the reference: 1 thread for evidence, one thread per
            mate file) on a bounded sample, timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_s_depleted"
UNIT = "reads/s"
WORKLOAD = "classifier: synthetic 10M 2x150 pairs + Kraken2 reads/report, -T Chordata -D 9606, deplete"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML DURING the timed region (a thread polling every
    millisecond: the timed region of the device arm is only tens of milliseconds long)"""

    def __init__(self, device: int):
        self.device = device
        self.sm, self.reasons = [], set()
        self.mx = None
        self._stop = threading.Event()
        self.th = None
        self.err = None

    def _run(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown,
                "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown,
                "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        while True:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                return
            if self._stop.wait(0.001):
                return

    def start(self):
        try:
            import pynvml as nv

            nv.nvmlInit()
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            idx = self.device
            if vis and all(x.strip().isdigit() for x in vis.split(",")):
                idx = int(vis.split(",")[self.device])
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._run, args=(nv, h), daemon=True)
            self.th.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def stop(self) -> dict:
        self._stop.set()
        if self.th:
            self.th.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.mx, "reasons": [f"nvml unavailable: {self.err}"], "samples": 0}
        return {"sm_mhz": statistics.median(self.sm), "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def bind_to_gpu_numa_node(device: int):
    """Best effort: run this rank on the CPUs next to its GPU (sysfs local_cpulist of the PCI device) so that
    the pinned host buffers of the end-to-end arm are allocated on the GPU's NUMA node.  Returns the cpulist."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = device
        if vis and all(x.strip().isdigit() for x in vis.split(",")):
            idx = int(vis.split(",")[device])
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(idx)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}/local_cpulist"
        cpulist = open(path).read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return cpulist
    except Exception as e:  # noqa: BLE001
        return f"unbound ({type(e).__name__})"


def taxids_for_config():
    """host stage: report -> taxid strings (-T Chordata -D 9606), the C++ state machine"""
    from scrubby_b200 import hostlib, synth

    return hostlib.get_taxids_from_report(synth.gen_kraken_report(5000), ["Chordata"], ["9606"])


# ------------------------------------------------------------------------------------------------ reference arm
_CPU_SAMPLE = {}


def cpu_sample_run(pairs: int):
    """the CPU oracle on `pairs` pairs: evidence -> set on 1 thread, the two mate files on 2 threads
    (cleaner.rs:238-248).  Inputs are generated once (untimed).  Returns (seconds, reads, fastq_bytes)."""
    from oracle import oracle as orc
    from scrubby_b200 import synth

    if pairs not in _CPU_SAMPLE:
        _CPU_SAMPLE[pairs] = (synth.gen_kraken_report(5000), synth.gen_kraken_reads(pairs).numpy(),
                              [synth.gen_fastq(pairs, m).numpy() for m in (1, 2)])
    rep, kr, fq = _CPU_SAMPLE[pairs]
    orc.lib()
    t0 = time.perf_counter()
    tax = orc.taxids_from_report(rep, ["Chordata"], ["9606"])
    ids = orc.set_from_reads(kr, 0, tax)
    res = [None, None]

    def one(i):
        res[i] = orc.clean_fastq(fq[i], ids, False, want_bytes=False)

    th = [threading.Thread(target=one, args=(i,)) for i in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    dt = time.perf_counter() - t0
    reads = res[0].reads_in + res[1].reads_in
    return dt, reads, int(fq[0].size + fq[1].size)


def cpu_sample_loop(pairs: int, min_seconds: float, max_reps: int = 200):
    """repeats the bounded sample until `min_seconds` of CPU work have been timed; returns the totals"""
    tot_t, tot_reads, tot_bytes, reps = 0.0, 0, 0, 0
    while reps < max_reps and (tot_t < min_seconds or reps == 0):
        dt, reads, nbytes = cpu_sample_run(pairs)
        tot_t += dt
        tot_reads += reads
        tot_bytes += nbytes
        reps += 1
    return tot_t, tot_reads, tot_bytes, reps


def fused_traffic(pairs: int, split: bool):
    """DRAM bytes per launch of the fused kernel from the committed `ncu --set full` capture of this workload
    (profiles/fused_traffic.json, written by tools/ncu_summary.py); None when no capture matches."""
    p = os.path.join(ROOT, "profiles", "fused_traffic.json")
    try:
        with open(p) as f:
            j = json.load(f)
        if j.get("pairs") == pairs and bool(j.get("split")) == bool(split):
            return j["dram_bytes_read"] + j["dram_bytes_write"]
    except Exception:
        pass
    return None


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.cpu_pairs
    times = []
    reps = 1
    for i in range(args.warmup + args.steps):
        # one step = the bounded sample repeated for about args.cpu_step_seconds of CPU work
        dt, reads, nbytes, reps = cpu_sample_loop(sample, args.cpu_step_seconds if i >= args.warmup else 0.0)
        if i >= args.warmup:
            times.append(dt)
    t = sum(times) / len(times)
    value = reads / t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample_pairs_per_step": sample * reps,
                   "outputs": "kept (reference-equivalent single output)", "fastq_gb_per_s": nbytes / t / 1e9},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 2, "kind": "port",
                         "sample": f"{reps} x ({sample} pairs + {sample} Kraken2 lines) = {nbytes / 1e9:.2f} GB FASTQ per step; "
                                   "C oracle structured like the reference (evidence on 1 thread, one thread per "
                                   "mate file); the Rust reference itself cannot be built here (no cargo/rustc)",
                         "cores_available": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from scrubby_b200 import api, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cpulist = bind_to_gpu_numa_node(local) if not args.no_bind else "unbound (--no-bind)"
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pairs = args.pairs
    start = rank * pairs  # weak scaling: every rank owns its own shard of reads and of the Kraken2 lines
    taxids = taxids_for_config()
    ctx = api.Context(local)  # enqueues on torch's current stream
    d_r = [synth.gen_fastq(pairs, m, device=dev, start=start) for m in (1, 2)]
    d_k = synth.gen_kraken_reads(pairs, device=dev, start=start)
    n_r = [int(t.numel()) for t in d_r]
    n_k = int(d_k.numel())
    d_out = [torch.empty(n + 64, dtype=torch.uint8, device=dev) for n in n_r]
    d_oth = [torch.empty(n + 64, dtype=torch.uint8, device=dev) for n in n_r] if args.split else [None, None]
    torch.cuda.synchronize()

    def step_dev():
        ids = api.IdSet.from_reads(ctx, d_k, 0, taxids)
        r = [api.clean_fastq_dev(ctx, ids, d_r[i], n_r[i], d_out[i], d_oth[i]) for i in range(2)]
        ids.free()
        return r

    # ---- device-resident arm
    ctx.set_profiling(False)
    for _ in range(args.warmup):
        res = step_dev()
    assert all(r.path == 1 for r in res), "the fused kernel must be the path that runs"
    ctx.set_profiling(True)
    ctx.fused_stats()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = ctx.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # `ncu --profile-from-start off` lists exactly the timed launches
    e0.record()
    for _ in range(args.steps):
        res = step_dev()
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (ctx.launches - l0)
    f_ms, f_n, f_bytes = ctx.fused_stats()
    ctx.set_profiling(False)
    reads_step = sum(r.reads_in for r in res)
    kept_step = sum(r.reads_out for r in res)

    # ---- end-to-end arm: pinned host buffers through the host-pointer C ABI (H2D + D2H inside)
    h_r = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in n_r]
    h_k = torch.empty(n_k, dtype=torch.uint8).pin_memory()
    for i in range(2):
        h_r[i].copy_(d_r[i])
    h_k.copy_(d_k)
    h_out = [torch.empty(n + 64, dtype=torch.uint8).pin_memory() for n in n_r]
    h_oth = [torch.empty(n + 64, dtype=torch.uint8).pin_memory() for n in n_r] if args.split else [None, None]
    torch.cuda.synchronize()

    def step_host():
        ids = api.IdSet.from_reads(ctx, h_k, 0, taxids)
        r = [api.clean_fastq_host(ctx, ids, h_r[i], n_r[i], h_out[i], h_oth[i]) for i in range(2)]
        ids.free()
        return r

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # warm-up: first touches of the pinned buffers, the chunk buffers of the pipeline, the memory pool.  The host side
    # settles by TIME rather than by step count (the first seconds after pinning GBs of host memory show sporadic
    # 0.3-0.6 s stalls): at least three steps, then until a step is within 15 % of the fastest seen, at most twelve
    warm_each = []
    while len(warm_each) < 12:
        t0 = time.perf_counter()
        rh = step_host()
        torch.cuda.synchronize()
        warm_each.append(time.perf_counter() - t0)
        if len(warm_each) >= 3 and warm_each[-1] <= 1.15 * min(warm_each) and warm_each[-2] <= 1.15 * min(warm_each):
            break
    if world > 1:  # every rank leaves the warm-up together
        w_t = torch.tensor([len(warm_each)], dtype=torch.int64, device=dev)
        dist.all_reduce(w_t, op=dist.ReduceOp.MAX)
        for _ in range(int(w_t) - len(warm_each)):
            rh = step_host()
    barrier()
    e2e_each = []
    for _ in range(e2e_steps):
        t0 = time.perf_counter()
        rh = step_host()
        torch.cuda.synchronize()
        e2e_each.append(time.perf_counter() - t0)
    t_e2e = sum(e2e_each) / e2e_steps
    d2h = sum(r.n_written + r.n_other for r in rh)
    h2d = sum(n_r) + n_k
    # cheap parity guard on the bench data itself: device and host arms agree, counts add up
    assert [r.n_written for r in rh] == [r.n_written for r in res]
    assert reads_step == 2 * pairs

    # ---- max over ranks
    ms_t = torch.tensor([ms, t_e2e * 1e3], dtype=torch.float64, device=dev)
    cnt = torch.tensor([reads_step, kept_step, sum(n_r)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)  # report counters, NCCL allreduce
    ms, e2e_ms = float(ms_t[0]), float(ms_t[1])
    reads_all, kept_all, bytes_all = (int(x) for x in cnt)

    if rank == 0:
        peak, peak_src = peaks()
        achieved = f_bytes / (f_ms * 1e-3) / 1e9 if f_ms > 0 else 0.0
        cpu = None
        if world == 1 or True:
            cpu_sample_run(args.cpu_pairs)  # warm the page cache / allocator
            dt, creads, cbytes, reps = cpu_sample_loop(args.cpu_pairs, args.cpu_seconds)
            cpu = {"value": creads / dt, "unit": UNIT, "cores": 2, "kind": "port",
                   "sample": f"{reps} x ({args.cpu_pairs} pairs + {args.cpu_pairs} Kraken2 lines) = "
                             f"{cbytes / 1e9:.2f} GB FASTQ, {dt:.2f} s; C oracle structured like the reference (1 thread evidence, 1 thread per "
                             "mate file)", "cores_available": os.cpu_count(),
                   "fastq_gb_per_s": cbytes / dt / 1e9}
        line = {
            "metric": METRIC, "value": reads_all / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {
                "workload": WORKLOAD if pairs == 10_000_000 else f"classifier: synthetic {pairs} 2x150 pairs + Kraken2 reads/report",
                "pairs_per_gpu": pairs, "fastq_bytes_per_gpu": sum(n_r), "kraken_bytes_per_gpu": n_k,
                "outputs": "kept+removed" if args.split else "kept (reference-equivalent single output)",
                "fraction_kept": kept_all / reads_all, "parallelism": f"chunk-sharded x{world}",
                "host_cpus_rank0": cpulist,
                "l2": "inputs (6.6 GB per GPU) are far larger than the 126 MB L2; no explicit flush",
                "fastq_gb_per_s": bytes_all / (ms * 1e-3) / 1e9,
            },
            "e2e": {"value": reads_all / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "ms_each_rank0": [round(x * 1e3, 2) for x in e2e_each],
                    "warmup_ms_each_rank0": [round(x * 1e3, 2) for x in warm_each],
                    "timing": "host wall clock around sgpu_idset_from_reads + 2x sgpu_clean_fastq on pinned host "
                              "buffers, stream synchronised"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": fused_traffic(pairs, args.split),
                         "kernel": "fastq_fused_kernel", "launches": f_n, "avg_ms": f_ms / f_n if f_n else None,
                         "algorithmic_bytes_per_launch": f_bytes / f_n if f_n else None, "peak_source": peak_src,
                         "share_of_step": f_ms / (ms * args.steps) if ms else None},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=10_000_000, help="pairs per GPU (10M = BASELINE configs[1])")
    ap.add_argument("--cpu-pairs", type=int, default=1_000_000, help="bounded CPU-baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU work timed for cpu_baseline")
    ap.add_argument("--cpu-step-seconds", type=float, default=3.0, help="--impl reference: CPU work per step")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--split", action="store_true", help="also write the removed records (kept + removed)")
    ap.add_argument("--no-bind", action="store_true", help="do not bind the rank to its GPU's NUMA-local CPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
