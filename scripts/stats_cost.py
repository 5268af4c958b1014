import time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, 100_000_000, 64, dev); q = engine.synth_codes(4001, 0, 8192, 64, dev)
idx = HammingIndex(db, 0, nd_total=db.n)
for _ in range(3): idx.search_packed(q, 1000)
torch.cuda.synchronize()
def run(label, mk):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    st = mk()
    t0 = time.perf_counter(); e0.record()
    for _ in range(5): keys = idx.search_packed(q, 1000, stats=st)
    e1.record(); torch.cuda.synchronize()
    print(label, "wall", round((time.perf_counter() - t0) / 5 * 1e3, 2), "events", round(e0.elapsed_time(e1) / 5, 2))
run("no stats", lambda: None)
run("stats {}", lambda: {})
run("time_collect", lambda: {"time_collect": True})
run("no stats", lambda: None)
