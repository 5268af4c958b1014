import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine, calc_utils as cu
from cmh_b200.index import HammingIndex
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
Q, D, K = 8192, 100_000_000, 1000
db = engine.synth_codes(4000, 0, D, 64, dev)
q = engine.synth_codes(4001, 0, Q, 64, dev)
db_host = torch.empty((D, 1), dtype=torch.int64).pin_memory(); db_host.copy_(db.sign)
db_dev = torch.empty_like(db.sign)
print("pinned:", db_host.is_pinned(), db_host[5:100].is_pinned(), db_host.view(torch.int64)[1000:2000].is_pinned())
for rep in range(3):
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t0.record()
    idx = HammingIndex.from_packed_host(db_host, 64, 0, nd_total=D, out=db_dev, pieces=3)
    # replace the copy events by timing events? they are not timing-enabled; add our own markers on the copy stream instead
    st = {"time_collect": True}
    k = idx.search_packed(q, K, stats=st)
    t1 = torch.cuda.Event(enable_timing=True); t1.record()
    torch.cuda.synchronize()
    ev = st["collect_events"]
    print("total", round(t0.elapsed_time(t1), 2), "launch windows:", [(round(t0.elapsed_time(ev[i]), 2), round(t0.elapsed_time(ev[i + 1]), 2)) for i in range(0, len(ev), 2)])
