"""Sweep the cut points of the prefix rule of the tensor-core top-K (run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
Q, D, K = 8192, int(os.environ.get("D", 100_000_000)), 1000
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, 64, dev); q = engine.synth_codes(4001, 0, Q, 64, dev)
stride = max(1, D // 65536)
sample = engine.PackedSet(db.sign[::stride].contiguous(), None, None, (D + stride - 1) // stride, 64)
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
for fr in ((), (0.5,), (0.5, 0.75), (0.33, 0.66), (0.25, 0.5, 0.75), (0.4, 0.6, 0.8), (0.3, 0.5, 0.7, 0.85), (0.2, 0.4, 0.6, 0.8), (0.15, 0.3, 0.45, 0.6, 0.8)):
    engine.TC_PREFIX_FRACTIONS = fr
    st = {}; buf = {}
    t = timed(lambda: engine.topk_tc(q, db, K, 0, sample=sample, stats=st, buffers=buf))
    print(json.dumps({"fractions": fr, "ms": round(t, 2), "cand_mean": round(float(st["candidates"].float().mean()), 1), "n_fail": st["n_fail"]}), flush=True)
