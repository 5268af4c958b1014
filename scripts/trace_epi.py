"""Epilogue timeline of tc_collect_kernel from the pipeline trace (needs a -DCMH_TC_TRACE build): per group, cycles
spent waiting for the accumulator tile, loading it, scanning (+ hit path), for a uniform threshold THR."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from cmh_b200 import _cabi, engine
Q, D, BITS = 8192, 25_000_000, 64
dev = torch.device("cuda", 0)
L = ctypes.CDLL(_cabi.LIB_PATH)
engine._cabi.lib()
db = engine.synth_codes(4000, 0, D, BITS, dev); q = engine.synth_codes(4001, 0, Q, BITS, dev)
b = engine.TcBuffers(Q, [D], BITS, 32768, dev)
st = engine._stream(dev); p = engine._ptr
ITERS, EV, ROLES = 96, 8, 6
for thr in [int(x) for x in sys.argv[1:]] or [-1, 15, 16]:
    thr0 = torch.full((Q,), thr, dtype=torch.int32, device=dev)
    tr = torch.zeros(ROLES * ITERS * EV, dtype=torch.int64, device=dev)
    L.cmh_tc_set_trace(ctypes.c_void_p(tr.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    engine.check(_cabi.lib().cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(thr0), b.seg_total, b.seg_cap, p(b.cand), p(b.cnt), p(b.aux), 0, st))
    e1.record(); torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(ROLES, ITERS, EV)
    print(f"=== thr {thr}: launch {e0.elapsed_time(e1):.2f} ms, candidates/query {float(b.cnt.sum(0).float().mean()):.0f}")
    for g in range(4):
        e = t[2 + g, 8:90]                     # rounds 8..89: [pre t_full, post t_full, loads done, -, scans done]
        wait = e[:, 1] - e[:, 0]; load = e[:, 2] - e[:, 1]; scan = e[:, 4] - e[:, 2]; cyc = e[1:, 0] - e[:-1, 0]
        print(f"  group {g}: cycle {cyc.mean():.0f} (p90 {np.percentile(cyc, 90):.0f})  wait {wait.mean():.0f}  load {load.mean():.0f}  scan+hits {scan.mean():.0f} (p90 {np.percentile(scan, 90):.0f}, max {scan.max()})")
    i = t[0, 8:90]                             # issuers: [post b_full, post t_empty wait, post commit, loop top]
    print(f"  issuers: wait b_full {np.mean(i[:, 0] - i[:, 3]):.0f}  wait t_empty {np.mean(i[:, 1] - i[:, 0]):.0f}  issue {np.mean(i[:, 2] - i[:, 1]):.0f}")
    pr = t[1, 8:90]                            # producer: [top, r_full ok, pre b_empty, post b_empty, stores done, fence done, arrived]
    print(f"  producer: cycle {np.mean(pr[1:, 0] - pr[:-1, 0]):.0f}  wait ring {np.mean(pr[:, 1] - pr[:, 0]):.0f}  wait b_empty {np.mean(pr[:, 3] - pr[:, 2]):.0f}  expand {np.mean(pr[:, 4] - pr[:, 3]):.0f}  fence {np.mean(pr[:, 5] - pr[:, 4]):.0f}")
