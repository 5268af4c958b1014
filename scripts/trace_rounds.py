"""Round-by-round timeline of one accumulator buffer of tc_collect_kernel (needs a -DCMH_TC_TRACE build): the issuer's
and the epilogue group's events of CTA (0,0), rounds 20..60, for a uniform threshold THR (argv[1])."""
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
thr = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr0 = torch.full((Q,), thr, dtype=torch.int32, device=dev)
tr = torch.zeros(ROLES * ITERS * EV, dtype=torch.int64, device=dev)
L.cmh_tc_set_trace(ctypes.c_void_p(tr.data_ptr()))
engine.check(_cabi.lib().cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(thr0), b.seg_total, b.seg_cap, p(b.cand), p(b.cnt), p(b.aux), 0, st))
torch.cuda.synchronize()
t = tr.cpu().numpy().reshape(ROLES, ITERS, EV)
t0 = t[t > 0].min()
print(f"thr {thr} group/buffer {g}: per round r (iteration it = 4 r + g):")
print("  issuer: b_full ok | t_empty ok | committed      epilogue (warp of lane 0): loop top | t_full seen | loaded+released | scan/park done")
prev = None
for r in range(20, 60):
    it = 4 * r + g
    if it >= ITERS: break
    i = t[0, it, :4] - t0; e = t[2 + g, r, :5] - t0
    print(f"  r {r:2d}  issuer {i[0]:7d} {i[1]:7d} {i[2]:7d}   epi {e[0]:7d} {e[1]:7d} {e[2]:7d} {e[4]:7d}   "
          f"[t_empty wait {i[1]-i[0]:5d} issue {i[2]-i[1]:5d} | mma->seen {e[1]-i[2]:5d} ld {e[2]-e[1]:4d} scan {e[4]-e[2]:5d} top->seen {e[1]-e[0]:5d}]")
