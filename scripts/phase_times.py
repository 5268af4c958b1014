"""Device-side phase timing of HammingIndex.search_packed (the library's own CUDA events, cmh_tc_timing)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
Q, D, K = int(os.environ.get("Q", 8192)), int(os.environ.get("D", 100_000_000)), int(os.environ.get("K", 1000))
BITS = int(os.environ.get("BITS", 64))
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, BITS, dev)
qs = [engine.synth_codes(4001, i * Q, Q, BITS, dev) for i in range(8)]
if os.environ.get("SAMPLE_ROWS"):
    HammingIndex.SAMPLE_ROWS = int(os.environ["SAMPLE_ROWS"])
idx = HammingIndex(db, 0, nd_total=D, assume_binary=True)
for i in range(3):
    idx.search_packed(qs[i], K)
torch.cuda.synchronize()
ts = []
for i in range(3, 8):
    torch.cuda.synchronize(); t0 = time.perf_counter(); idx.search_packed(qs[i], K); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("search_packed wall ms:", [round(t, 2) for t in ts])
st = {"time_phases": True}
for i in range(3, 8):
    idx.search_packed(qs[i], K, stats=st)
n = st["timed_searches"]
print("phases ms/search:", {k: round(v / n, 3) for k, v in st["phase_ms_sum"].items()}, "collect", round(st["collect_ms_sum"] / n, 3))
print("launches ms (last search):", [round(x, 3) for x in st["launch_ms"]], "n_fail", st["n_fail"], "candidates/query",
      float(st["candidates"].float().mean()))
