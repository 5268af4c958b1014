"""Pipeline trace of tc_collect_kernel (needs a -DCMH_TC_TRACE build of the library)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from cmh_b200 import _cabi, engine
Q, D, BITS = 8192, 25_000_000, 64
dev = torch.device("cuda", 0)
L = ctypes.CDLL(_cabi.LIB_PATH)
engine._cabi.lib()
db = engine.synth_codes(4000, 0, D, BITS, dev); q = engine.synth_codes(4001, 0, Q, BITS, dev)
b = engine.TcBuffers(Q, [D], BITS, 32768, dev)
thr0 = torch.full((Q,), -1, dtype=torch.int32, device=dev)
st = engine._stream(dev); p = engine._ptr
ITERS, EV, ROLES = 96, 8, 6
for mode in [int(m) for m in sys.argv[1:]] or [3, 0]:
    tr = torch.zeros(ROLES * ITERS * EV, dtype=torch.int64, device=dev)
    L.cmh_tc_set_trace(ctypes.c_void_p(tr.data_ptr()))
    engine.check(_cabi.lib().cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(thr0), b.seg_total, b.seg_cap, p(b.cand), p(b.cnt), p(b.aux), mode, st))
    torch.cuda.synchronize()
    t = tr.cpu().numpy().reshape(ROLES, ITERS, EV)
    t0 = t[t > 0].min()
    t = np.where(t > 0, t - t0, -1)
    print(f"=== mode {mode}")
    print("MMA issuers: it: [post-b_full/pre-t_empty-wait, post-wait, post-commit, loop top]")
    for it in list(range(0, 16)) + list(range(64, 96)):
        print("  it", it, "issuer", it % 4, t[0, it, :4].tolist())
    print("producer: i: [top, r_full ok, pre b_empty, post b_empty, stores done, fence done, arrived]")
    for i in list(range(0, 12)) + list(range(40, 60)):
        print("  i", i, t[1, i, :7].tolist())
    for g in range(4):
        print(f"epilogue group {g}: round: [pre t_full, post t_full, loads done, -, scans done]")
        for r in list(range(0, 4)) + list(range(16, 24)):
            print("  r", r, t[2 + g, r, :5].tolist())
