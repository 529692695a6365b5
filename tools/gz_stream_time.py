"""Wall time of the CLI on plain-gzip paired FASTQ with the inflate -> filter -> deflate pipeline (default) and with the
whole-file path (SCRUBBY_NO_GZ_STREAM=1); gz in, gz out and gz in, plain out.  python tools/gz_stream_time.py [pairs]"""
import os
import subprocess
import sys
import tempfile
import time
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scrubby_b200 import hostlib, synth  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
hostlib.build()
d = tempfile.mkdtemp(prefix="gzs")
t0 = time.time()
ins = []
for m in (1, 2):
    raw = synth.gen_fastq(pairs, m).numpy().tobytes()
    co = zlib.compressobj(1, zlib.DEFLATED, 31)
    p = os.path.join(d, f"r{m}.fq.gz")
    with open(p, "wb") as f:
        f.write(co.compress(raw) + co.flush())
    ins.append(p)
    print(f"mate {m}: {len(raw) / 1e6:.0f} MB -> {os.path.getsize(p) / 1e6:.0f} MB gz", flush=True)
ids = os.path.join(d, "ids.txt")
with open(ids, "wb") as f:
    f.write(synth.gen_txt_ids(pairs).numpy().tobytes())
print(f"setup {time.time() - t0:.1f} s", flush=True)
for outs in ((".fq.gz", "gz in -> gz out"), (".fq", "gz in -> plain out")):
    for stream in (True, False, True, False):
        env = dict(os.environ)
        if not stream:
            env["SCRUBBY_NO_GZ_STREAM"] = "1"
        o = [os.path.join(d, f"o{m}{outs[0]}") for m in (1, 2)]
        t = time.time()
        r = subprocess.run([hostlib.CLI, "alignment", "-i", *ins, "-o", *o, "-a", ids, "--format", "txt"], env=env,
                           capture_output=True, text=True)
        dt = time.time() - t
        assert r.returncode == 0, r.stderr[-500:]
        print(f"{outs[1]}, {'pipeline  ' if stream else 'whole file'}: {dt:.2f} s  ({2 * pairs / dt / 1e6:.2f} M reads/s)", flush=True)
