"""A/B of the hit path inside one process, interleaved so that clocks and temperature are shared: all launches on the
draining-warp kernel (0), all on the hit-worker kernel (1), per-launch choice (-1: pilot draining, main workers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import _cabi, engine
from cmh_b200.index import HammingIndex
Q, D, K = 8192, int(os.environ.get("D", 100_000_000)), 1000
dev = torch.device("cuda", 0)
db = engine.synth_codes(4000, 0, D, 64, dev)
qs = [engine.synth_codes(4001, i * Q, Q, 64, dev) for i in range(6)]
idx = HammingIndex(db, 0, nd_total=D, assume_binary=True)
L = _cabi.lib()
res = {m: [] for m in (0, 1, -1)}
launches = {m: [] for m in (0, 1, -1)}
for rep in range(int(os.environ.get("REPS", 8))):
    for m in (0, 1, -1):
        L.cmh_tc_set_workers(m)
        st = {"time_phases": True}
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); idx.search_packed(qs[rep % 6], K, stats=st); b.record(); torch.cuda.synchronize()
        if rep >= 2:
            res[m].append(a.elapsed_time(b)); launches[m].append(st["launch_ms"])
L.cmh_tc_set_workers(-1)
for m in (0, 1, -1):
    print("mode", m, "median ms/step", round(float(np.median(res[m])), 3), "min", round(min(res[m]), 3),
          "launch medians", [round(float(x), 3) for x in np.median(np.array(launches[m]), axis=0)])
