"""Warp instructions and stall samples summed over source-line ranges of one file.
python tools/ncu_regions.py rep file.cu name:lo-hi name:lo-hi ...   (lines outside every range -> 'other')"""
import csv, io, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
regions = []
for a in sys.argv[3:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
STALLS = ["barrier", "long_sb", "short_sb", "wait", "branch_resolving", "selected", "not_selected", "math", "mio", "lg", "no_inst", "dispatch", "sleep", "membar"]
cur = None; hdr = None
acc = {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; hdr = None; continue
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}; continue
    if hdr is None or len(r) < len(hdr): continue
    try: ln = int(r[0])
    except ValueError: continue
    def num(k):
        try: return int(r[hdr[k]])
        except (ValueError, KeyError): return 0
    name = "other:" + cur
    if cur == fname:
        name = "other"
        for n, lo, hi in regions:
            if lo <= ln <= hi: name = n; break
    a = acc.setdefault(name, [0, 0] + [0] * len(STALLS))
    a[0] += num("Instructions Executed"); a[1] += num("# Samples")
    for si, sn in enumerate(STALLS): a[2 + si] += num("stall_" + sn)
ti = sum(a[0] for a in acc.values()); ts = sum(a[1] for a in acc.values())
print(f"total warp-inst {ti} samples {ts}")
for n, a in sorted(acc.items(), key=lambda x: -x[1][0]):
    if a[0] == 0 and a[1] == 0: continue
    st = " ".join(f"{sn[:6]}={a[2+si]}" for si, sn in enumerate(STALLS) if a[2+si] * 50 > max(a[1], 1))
    print(f"{n:28s} inst {a[0]:10d} ({100*a[0]/ti:4.1f}%) samp {a[1]:6d} ({100*a[1]/ts:4.1f}%) | {st}")
