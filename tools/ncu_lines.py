"""Per CUDA source line: warp instructions executed and stall samples.  python tools/ncu_lines.py rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur = None
hdr = None
agg = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        hdr = None
        continue
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        first_source = r.index("Source")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    def num(k):
        try:
            return int(r[hdr[k]])
        except (ValueError, KeyError):
            return 0
    inst = num("Instructions Executed")
    if inst == 0 and num("# Samples") == 0:
        continue
    agg.append((cur, ln, r[first_source].strip()[:100], inst, num("# Samples"), num("stall_barrier"), num("stall_long_sb"),
                num("stall_short_sb"), num("stall_wait"), num("L1 Wavefronts Shared Excessive")))
tot_i = sum(a[3] for a in agg); tot_s = sum(a[4] for a in agg)
print(f"total warp-inst {tot_i}  samples {tot_s}")
print("== by instructions")
for a in sorted(agg, key=lambda a: -a[3])[:topn]:
    print(f"{a[0]:18s}:{a[1]:4d} inst {a[3]:10d} ({100*a[3]/tot_i:4.1f}%) samp {a[4]:6d} ({100*a[4]/max(tot_s,1):4.1f}%) bar {a[5]:5d} lsb {a[6]:5d} ssb {a[7]:5d} wait {a[8]:5d} shx {a[9]:8d} | {a[2]}")
print("== by samples")
for a in sorted(agg, key=lambda a: -a[4])[:topn]:
    print(f"{a[0]:18s}:{a[1]:4d} inst {a[3]:10d} ({100*a[3]/tot_i:4.1f}%) samp {a[4]:6d} ({100*a[4]/max(tot_s,1):4.1f}%) bar {a[5]:5d} lsb {a[6]:5d} ssb {a[7]:5d} wait {a[8]:5d} shx {a[9]:8d} | {a[2]}")
