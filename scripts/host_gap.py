"""Where does the bench's step lose time against a bare search loop?  Variants: stats on/off, NVML sampler on/off."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
import bench
Q, D, K = 8192, 100_000_000, 1000
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, 64, dev); q = engine.synth_codes(4001, 0, Q, 64, dev)
idx = HammingIndex(db, 0, nd_total=D)
for _ in range(3):
    idx.search_packed(q, K, stats={"time_collect": True, "time_phases": True})
torch.cuda.synchronize()

def run(label, use_stats, n=5):
    stats = {"time_collect": True, "time_phases": True} if use_stats else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        keys = idx.search_packed(q, K, stats=stats)
    e1.record(); torch.cuda.synchronize()
    msg = f"{label:28s} {e0.elapsed_time(e1) / n:.2f} ms/step"
    if use_stats:
        pe = stats["phase_events"]; ph = {}
        for (n0, a), (n1, b) in zip(pe, pe[1:]):
            if n1 != "start":
                ph[n1] = ph.get(n1, 0.0) + a.elapsed_time(b) / n
        msg += "  " + " ".join(f"{k}={v:.2f}" for k, v in ph.items())
    print(msg, flush=True)

run("bare", False); run("stats", True); run("bare", False)
s = bench.ClockSampler(0); s.start(); time.sleep(0.2)
run("sampler bare", False); run("sampler stats", True)
s.stop()
run("after sampler bare", False)
