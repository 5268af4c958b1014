"""Per-phase device times of the tensor-core search (cmh_tc_timing) by code length and database size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
dev = torch.device("cuda", 0)
Q, K = 8192, 1000
for bits, D in ((64, 20_000_000), (32, 20_000_000), (16, 20_000_000), (32, 100_000_000), (64, 100_000_000)):
    db = engine.synth_codes(7000 + bits, 0, D, bits, dev)
    qs = [engine.synth_codes(7001 + bits, i * Q, Q, bits, dev) for i in range(5)]
    idx = HammingIndex(db, 0, nd_total=D, assume_binary=True)
    st = {"time_phases": True, "time_collect": True}
    for i in range(5):
        if i == 2: st = {"time_phases": True, "time_collect": True}
        idx.search_packed(qs[i], K, stats=st)
    n = st["timed_searches"]
    print(bits, D, {k: round(v / n, 3) for k, v in st["phase_ms_sum"].items()}, "launch_ms", [round(x, 3) for x in st["launch_ms"]],
          "n_fail", st["n_fail"], "cand/query", round(float(st["candidates"].sum()) / Q), "pilot_rows", st["pilot_rows"], flush=True)
    del db, idx, qs; torch.cuda.empty_cache()
