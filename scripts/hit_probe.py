"""Cost of the hit path of tc_collect_kernel by parts: one launch over D rows with a uniform threshold, the hit path
cut short at successive points (cmh_tc_probe: 32 flagged slices not parked, 16 parked slices dropped, 8 decoded but
not stored, 0 everything) and the real launch with in-kernel tightening (cmh_tc_collect, K > 0)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import _cabi, engine
Q, D, BITS, K = 8192, int(os.environ.get("D", 50_000_000)), 64, 1000
dev = torch.device("cuda", 0)
L = _cabi.lib(); st = engine._stream(dev); p = engine._ptr
db = engine.synth_codes(4000, 0, D, BITS, dev); q = engine.synth_codes(4001, 0, Q, BITS, dev)
b = engine.TcBuffers(Q, [D], BITS, 32768, dev)

def timed(fn, reps=6):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts)

base = None
for thr in (-1, 15, 16, 17, -1, 15, 16):
    thr0 = torch.full((Q,), thr, dtype=torch.int32, device=dev)
    row = []
    for mode in (32, 16, 64, 0):
        t = timed(lambda: engine.check(L.cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(thr0), b.seg_total, b.seg_cap, p(b.cand), p(b.cnt), p(b.aux), mode, st)))
        row.append(t)
    cand = float(b.cnt.sum(0).float().mean())
    def real():
        b.aux.zero_()
        engine.check(L.cmh_tc_collect(p(q.sign), Q, p(db.sign), D, BITS, 0, p(thr0), K, 0, b.seg_total, b.seg_cap, p(b.cand), p(b.cnt), p(b.aux), st))
    t_real = timed(real); cand_real = float(b.cnt.sum(0).float().mean())
    if base is None: base = row[-1]
    cyc = lambda t, c: (t - base) * 1e-3 * 148 * 1.965e9 / max(1.0, c * Q)
    print(f"thr {thr:3d}: no-park {row[0]:.2f}  no-work {row[1]:.2f}  eager {row[2]:.2f}  full {row[3]:.2f} ms  cand/q {cand:.0f}  -> {cyc(row[3], cand):.0f} SM-cycles/candidate;"
          f"  K>0: {t_real:.2f} ms cand/q {cand_real:.0f}", flush=True)
