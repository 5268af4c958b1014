"""Sweep the pilot stages / sample size of the tensor-core top-K (run on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
Q, D, K = 8192, 100_000_000, 1000
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, 64, dev); q = engine.synth_codes(4001, 0, Q, 64, dev)
def timed(fn, reps=4):
    fn(); fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sorted(ts)[len(ts) // 2]
for stride in (1525, 3050):
    sample = engine.PackedSet(db.sign[::stride].contiguous(), None, None, (D + stride - 1) // stride, 64)
    for fracs in ((64,), (512, 64), (1024, 64), (512, 48), (256, 32), (64,)):
        pilot = [(D // f) // 256 * 256 for f in fracs]
        st = {}; buf = {}
        t, med = timed(lambda: engine.topk_tc(q, db, K, 0, sample=sample, pilot=pilot, stats=st, buffers=buf))
        print(json.dumps({"sample_rows": sample.n, "pilot_fracs": fracs, "ms_min": round(t, 2), "ms_med": round(med, 2),
                          "cand_mean": round(float(st["candidates"].float().mean()), 1), "n_fail": st["n_fail"]}), flush=True)
