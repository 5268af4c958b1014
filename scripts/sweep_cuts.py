"""Where to apply the exact prefix rule: one GPU (fractions of the rows) and one shard of 8 (lockstep stripe boundaries,
loopback transport).  Device time per search by CUDA events, 6 distinct query chunks each."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import _cabi, engine, sharded
Q, D, K = 8192, 100_000_000, 1000
dev = torch.device("cuda", 0)
L = _cabi.lib()
qs = [engine.synth_codes(4001, i * Q, Q, 64, dev) for i in range(8)]

def timed(fn):
    for i in range(2): fn(qs[i], None)
    torch.cuda.synchronize(); ts = []
    st = {"time_phases": True}
    for i in range(2, 8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(qs[i], st); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts)), st

if os.environ.get("MODE", "one") == "one":
    db = engine.synth_codes(4000, 0, D, 64, dev)
    smp = engine.PackedSet(db.sign[::D // 65536].contiguous(), None, None, len(range(0, D, D // 65536)), 64)
    for fr in [(0.3, 0.5, 0.7, 0.85), (0.1, 0.2, 0.3, 0.5, 0.7, 0.85), (0.15, 0.3, 0.5, 0.7, 0.85), (0.1, 0.2, 0.35, 0.55, 0.8),
               (0.08, 0.16, 0.3, 0.5, 0.75), (0.2, 0.4, 0.6, 0.8)]:
        buf = {}
        ms, st = timed(lambda q, s: engine.topk_tc(q, db, K, 0, sample=smp, buffers=buf, prefix_fractions=fr, stats=s))
        print("one GPU", fr, "ms", round(ms, 3), "launches", [round(x, 2) for x in st["launch_ms"]], "n_fail", st["n_fail"],
              "cand", round(float(st["candidates"].float().mean())), flush=True)
else:
    W = int(os.environ.get("WORLD", 8))
    ptr = ctypes.POINTER(_cabi.Comm)()
    _cabi.check(L.cmh_comm_create_loopback(W, 0, ctypes.byref(ptr)), "loopback")
    class Loop:
        world, rank = W, 0
        def handle(self): return ptr
    for fr in [(0.3, 0.6), (0.1, 0.2, 0.4, 0.7), (0.15, 0.35, 0.65), (0.1, 0.25, 0.5, 0.75), (0.12, 0.3, 0.6), (0.2, 0.5)]:
        ranges, stripes = sharded.lockstep_stripes(D, W, 0, fractions=fr)
        rows = torch.cat([engine.synth_codes(4000, a, b - a, 64, dev).sign for a, b in ranges])
        db = engine.PackedSet(rows, None, None, rows.shape[0], 64)
        share = max(4096, 65536 * db.n // D)
        s_rows = db.sign[::max(1, db.n // share)].contiguous()
        smp = engine.PackedSet(s_rows, None, None, s_rows.shape[0], 64)
        buf = {}
        fb = lambda sub: torch.full((sub.n, K), -1, dtype=torch.int64, device=dev)
        ms, st = timed(lambda q, s: engine.topk_tc(q, db, K, ranges[0][0], sample=smp, comm=Loop(), nd_total=D, stripes=stripes,
                                                   buffers=buf, gather=False, stats=s, exact_fallback=fb))
        n = st["timed_searches"]
        ph = {k: round(v / n, 3) for k, v in st["phase_ms_sum"].items()}
        print(f"shard of {W}", fr, "device ms (phases)", round(sum(ph.values()), 3), ph, "launches", [round(x, 3) for x in st["launch_ms"]],
              "cand", round(float(st["candidates"].float().mean())), flush=True)
