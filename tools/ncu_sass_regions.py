"""Warp instructions / stall samples of one kernel by source region, counted per SASS instruction (each
address once, attributed to the innermost inlined source line) -- the CUDA-line rows of ncu's source page
count an inlined instruction at the callee line AND at every call site.
python tools/ncu_sass_regions.py rep file.cu name:lo-hi ...  [--lines N: also the top N source lines]"""
import csv, io, subprocess, sys
args = [a for a in sys.argv[1:] if not a.startswith("--lines")]
topn = 0
for a in sys.argv[1:]:
    if a.startswith("--lines="):
        topn = int(a.split("=")[1])
rep, fname = args[0], args[1]
regions = []
for a in args[2:]:
    n, r = a.split(":")
    lo, hi = r.split("-")
    regions.append((n, int(lo), int(hi)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
cur_file, cur_line = None, None
seen = {}  # address -> [inst, samples, candidates[(file, line)]]
for r in csv.reader(io.StringIO(out)):
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if not r or r[0] == "Line No" or r[0] == "Function Name":
        continue
    if r[0].isdigit():
        cur_line = int(r[0])
        continue
    if r[0] == "" and len(r) > 7 and r[2].startswith("0x"):
        try:
            inst, samp = int(r[7]), int(r[6])
        except ValueError:
            continue
        e = seen.setdefault(r[2], [inst, samp, []])
        e[2].append((cur_file, cur_line))
acc, lines = {}, {}
for addr, (inst, samp, cands) in seen.items():
    # innermost: a header / other file first, else the smallest line of the kernel's file (helpers come first)
    other = [c for c in cands if c[0] != fname]
    f, ln = other[0] if other else min(cands, key=lambda c: c[1])
    name = "other:" + f
    if f == fname:
        name = "other"
        for n, lo, hi in regions:
            if lo <= ln <= hi:
                name = n
                break
    a = acc.setdefault(name, [0, 0])
    a[0] += inst
    a[1] += samp
    l = lines.setdefault((f, ln), [0, 0])
    l[0] += inst
    l[1] += samp
ti = sum(a[0] for a in acc.values())
ts = sum(a[1] for a in acc.values())
print(f"total warp-inst {ti} samples {ts} ({len(seen)} SASS instructions)")
for n, a in sorted(acc.items(), key=lambda x: -x[1][0]):
    print(f"{n:32s} inst {a[0]:11d} ({100*a[0]/ti:4.1f}%)  samples {a[1]:7d} ({100*a[1]/max(ts,1):4.1f}%)")
if topn:
    print("-- top source lines by instructions")
    for (f, ln), a in sorted(lines.items(), key=lambda x: -x[1][0])[:topn]:
        print(f"{f}:{ln:5d} inst {a[0]:11d} ({100*a[0]/ti:4.1f}%) samples {a[1]:7d} ({100*a[1]/max(ts,1):4.1f}%)")
    print("-- top source lines by samples")
    for (f, ln), a in sorted(lines.items(), key=lambda x: -x[1][1])[:topn]:
        print(f"{f}:{ln:5d} inst {a[0]:11d} ({100*a[0]/ti:4.1f}%) samples {a[1]:7d} ({100*a[1]/max(ts,1):4.1f}%)")
