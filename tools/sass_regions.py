"""Static SASS instruction count per source-line region of the fused kernel.
python tools/sass_regions.py [lib.so]   (uses cuobjdump -xelf + nvdisasm -g)"""
import bisect, collections, os, re, subprocess, sys, tempfile
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "scrubby_b200", "lib", "libscrubby_gpu.so")
src = os.path.join(os.path.dirname(__file__), "..", "scrubby_b200", "csrc", "fastq_fused.cu")
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "fastq_fused", os.path.abspath(so)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(d, cub)], capture_output=True, text=True).stdout
# region marks from '// ----' comments and function heads in the source
marks = [(1, "top")]
for i, l in enumerate(open(src), 1):
    m = re.match(r"\s*// -{4,} (.*)", l)
    if m: marks.append((i, m.group(1)[:40]))
    m = re.match(r"\s*// ---- (\S+)", l)
    if m: marks.append((i, m.group(1)[:40]))
marks.sort()
infn = False; cur = None
reg = collections.Counter(); files = collections.Counter()
for line in dis.splitlines():
    if line.startswith(".text."):
        infn = "fastq_fused_kernel" in line; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+\S", line) and cur:
        if cur[0] == "fastq_fused.cu":
            i = bisect.bisect_right([x[0] for x in marks], cur[1]) - 1
            reg[f"{marks[i][0]:4d} {marks[i][1]}"] += 1
        else:
            reg[cur[0]] += 1
tot = sum(reg.values())
print("total SASS instructions", tot, "=", tot * 16 // 1024, "KiB")
for k, v in sorted(reg.items(), key=lambda kv: -kv[1]): print(f"{v:6d}  {k}")
