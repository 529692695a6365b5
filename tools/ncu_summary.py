"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__throughput.avg.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "launch__block_size"]
if len(sys.argv) > 4 and sys.argv[4] in ("c4", "c2"):  # python tools/ncu_summary.py rep top_n traffic.json c4|c2 n_gpus [split]
    import json

    v = rows[2]
    def _bytes(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(v[i].replace(",", "")) * scale
    with open(sys.argv[3], "w") as f:
        json.dump({"kernel": v[hdr.index("Kernel Name")], "config": sys.argv[4], "n_gpus": int(sys.argv[5]),
                   "split": len(sys.argv) > 6, "dram_bytes_read": _bytes("dram__bytes_read.sum"),
                   "dram_bytes_write": _bytes("dram__bytes_write.sum"),
                   "gpu_time_us_under_ncu": v[hdr.index("gpu__time_duration.sum")], "source": rep}, f, indent=1)
elif len(sys.argv) > 4:  # (round 1 form) python tools/ncu_summary.py rep top_n traffic.json pairs [split]
    import json

    v = rows[2]
    def _bytes(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(v[i].replace(",", "")) * scale
    with open(sys.argv[3], "w") as f:
        json.dump({"kernel": v[hdr.index("Kernel Name")], "pairs": int(sys.argv[4]), "split": len(sys.argv) > 5,
                   "dram_bytes_read": _bytes("dram__bytes_read.sum"), "dram_bytes_write": _bytes("dram__bytes_write.sum"),
                   "gpu_time_us_under_ncu": v[hdr.index("gpu__time_duration.sum")], "source": rep}, f, indent=1)
for vals in rows[2:]:
    print("== kernel", vals[hdr.index("Kernel Name")], "grid", vals[hdr.index("Grid Size")])
    for i, h in enumerate(hdr):
        if any(h == k or h.startswith(k) for k in keys) and "Triage" not in h and "per_second" not in h and \
                "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "").replace("sm__throughput.avg.pct_of_peak_sustained_elapsed", ""):
            print(f"  {h:80s} {units[i]:10s} {vals[i]}")
    print("  -- stall reasons (warps per issue-active cycle)")
    for i, h in enumerate(hdr):
        if "issue_stalled" in h and h.endswith(".ratio") and "not_issued" not in h:
            try:
                if float(vals[i]) >= 0.05:
                    print(f"  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s} {vals[i]}")
            except ValueError:
                pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
# locate header row
for hi, r in enumerate(rows):
    if "Instructions Executed" in r:
        break
h = rows[hi]
ci = {n: i for i, n in enumerate(h)}
data = [r for r in rows[hi + 1:] if len(r) == len(h)]
def num(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(num(r[ci["Instructions Executed"]]) for r in data)
samp = sum(num(r[ci["# Samples"]]) for r in data)
print(f"== source page: {len(data)} rows, {tot} warp-instructions, {samp} samples")
key = "Source" if "Source" in ci else h[1]
top = sorted(range(len(data)), key=lambda i: -num(data[i][ci["# Samples"]]))[:topn]
print("  top rows by stall samples: idx | samples | inst executed | avg threads | source")
for i in sorted(top):
    r = data[i]
    print(f"  {i:5d} {r[ci['# Samples']]:>7s} {r[ci['Instructions Executed']]:>11s} {r[ci['Avg. Threads Executed']]:>4s}  {r[ci[key]].strip()[:110]}")
