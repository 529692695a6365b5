import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from scrubby_b200 import api, synth
ctx = api.Context(0); dev = torch.device("cuda", 0)
paf = synth.gen_paf(1_000_000, device=dev)
for mode in (0, 1):
    ctx.set_mode(mode)
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ids = api.IdSet.from_paf(ctx, paf, 50, 0.5, 50); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1)); n = len(ids); ids.free()
    print("mode", mode, "from_paf ms", round(best, 3), "ids", n)
