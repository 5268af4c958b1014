"""Host-side phase timing of HammingIndex.search_packed (perf_counter around synchronised phases)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
Q, D, K = int(os.environ.get("Q", 8192)), int(os.environ.get("D", 100_000_000)), 1000
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, 64, dev); q = engine.synth_codes(4001, 0, Q, 64, dev)
idx = HammingIndex(db, 0, nd_total=D)
for _ in range(3):
    idx.search_packed(q, K)
torch.cuda.synchronize()
# whole call, async
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); idx.search_packed(q, K); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("search_packed wall ms:", [round(t, 2) for t in ts])
# monkeypatch ctypes entry points to sync + time
from cmh_b200 import _cabi
L = _cabi.lib()
acc = {}
class Timed:
    def __init__(self, name, fn): self.name, self.fn = name, fn
    def __call__(self, *a):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = self.fn(*a); torch.cuda.synchronize()
        acc.setdefault(self.name, []).append((time.perf_counter() - t0) * 1e3); return r
for n in ("cmh_eval_hist", "cmh_topk_threshold", "cmh_tc_collect", "cmh_tc_cand_hist", "cmh_tc_choose", "cmh_topk_finalize", "cmh_tc_plan", "cmh_eval_plan_design"):
    setattr(L, n, Timed(n, getattr(L, n)))
torch.cuda.synchronize(); t0 = time.perf_counter(); idx.search_packed(q, K); torch.cuda.synchronize()
total = (time.perf_counter() - t0) * 1e3
print("synchronised total ms:", round(total, 2))
for k, v in acc.items(): print(f"  {k:24s} {[round(x, 3) for x in v]}")
print("  unaccounted:", round(total - sum(sum(v) for v in acc.values()), 2))
