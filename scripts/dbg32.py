import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
dev = torch.device("cuda", 0)
for bits, nq, nd, K in ((32, 64, 300_000, 100), (16, 40, 2_000_000, 1000), (32, 64, 1_500_000, 1000)):
    db = engine.synth_codes(300 + bits, 0, nd, bits, dev)
    q = engine.synth_codes(400 + bits, 0, nq, bits, dev)
    want = engine.RankPass(q, db, need_labels=False).topk(K, 3)
    st = {}
    got = engine.topk_tc(q, db, K, 3, stats=st)
    bad = (got != want)
    print(bits, nd, K, "equal", bool(torch.equal(got, want)), "n_fail", st["n_fail"], "bad queries", int(bad.any(1).sum()), "bad entries", int(bad.sum()),
          "thr", st["thr"][:8].tolist(), "thr_final", st["thr_final"][:8].tolist(), "cand", int(st["candidates"][:8].sum()))
    if bad.any():
        qi = int(bad.any(1).nonzero()[0]); pos = int(bad[qi].nonzero()[0])
        g, w = got[qi].tolist(), want[qi].tolist()
        print(" query", qi, "first bad pos", pos, "got", [(k >> 32, k & 0xffffffff) for k in g[max(0,pos-2):pos+4]], "want", [(k >> 32, k & 0xffffffff) for k in w[max(0,pos-2):pos+4]])
        sg, sw = set(g), set(w)
        print(" missing from got:", [(k >> 32, k & 0xffffffff) for k in sorted(sw - sg)[:6]], " extra in got:", [(k >> 32, k & 0xffffffff) for k in sorted(sg - sw)[:6]])
