"""Per-tile timeline statistics of the fused kernel from a -DSGPU_FUSED_TIMING build's trace
(SGPU_FUSED_TRACE=file: 8 u64 per tile = load issued, loaded, aggregate published, inclusive published, copy done
[globaltimer ns], smid).   python tools/trace_stats.py trace.bin"""
import sys

import numpy as np

tr = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
ok = (tr[:, :5] > 0).all(axis=1)
print(f"{len(tr)} tiles, {ok.sum()} fully traced")
t0 = tr[ok, 0].min()
issue, loaded, agg, inc, done, sm = [tr[:, i] - (t0 if i < 5 else 0) for i in range(6)]
span = (done[ok].max() - issue[ok].min()) / 1e3
print(f"kernel span {span:.1f} us, {len(tr) / span:.1f} tiles/us")


def q(name, x):
    x = x[ok] / 1e3
    print(f"{name:34s} mean {x.mean():6.2f}  p10 {np.percentile(x, 10):6.2f}  p50 {np.percentile(x, 50):6.2f}  "
          f"p90 {np.percentile(x, 90):6.2f}  p99 {np.percentile(x, 99):6.2f}  max {x.max():6.2f} us")


q("load (issue -> loaded)", loaded - issue)
q("parse+probe (loaded -> aggregate)", agg - loaded)
q("in-order wait (agg -> inclusive)", inc - agg)
q("copy-out (inclusive -> done)", done - inc)
q("residency (issue -> done)", done - issue)
# the in-order wait split: until every earlier aggregate is there / the look-back chain after that
pm = np.maximum.accumulate(agg)
prev = np.concatenate([[0], pm[:-1]])
ready = np.maximum(agg, prev)
q("  wait for earlier aggregates", ready - agg)
q("  look-back chain after that", inc - ready)
# who holds the others back: tiles whose aggregate raises the running maximum, by how much
rise = agg - prev
block = ok & (rise > 0)
print(f"tiles that raise the frontier of aggregates: {block.sum()} ({100 * block.sum() / ok.sum():.1f} %)")
lead = (agg - loaded)[ok].mean()
for name, sel in (("frontier tiles", block), ("all tiles", ok)):
    print(f"  {name:16s} load {((loaded - issue)[sel]).mean() / 1e3:5.2f}  parse {((agg - loaded)[sel]).mean() / 1e3:5.2f} us")
# systematic differences between SMs
sms = np.unique(sm[ok])
per = np.array([[(loaded - issue)[ok & (sm == s)].mean(), (agg - loaded)[ok & (sm == s)].mean(),
                 (done - inc)[ok & (sm == s)].mean(), (ok & (sm == s)).sum(), (block & (sm == s)).sum()] for s in sms])
for j, name in enumerate(("load", "parse", "copy")):
    c = per[:, j] / 1e3
    print(f"per-SM mean {name:6s}: min {c.min():5.2f}  p50 {np.median(c):5.2f}  max {c.max():5.2f} us  (SM {sms[c.argmax()]} slowest)")
print("tiles per SM: min", int(per[:, 3].min()), "max", int(per[:, 3].max()))
share = per[:, 4] / np.maximum(per[:, 3], 1)
o = np.argsort(-share)[:8]
print("SMs with the largest share of frontier tiles:", [(int(sms[i]), round(float(share[i]), 2)) for i in o])
# distance to the nearest inclusive descriptor when the look-back could start
order = np.argsort(inc)
