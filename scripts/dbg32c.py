"""Bare cmh_tc_collect launches against brute force under several variations (see the prints)."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import engine, _cabi
dev = torch.device("cuda", 0)
L = _cabi.lib()
def run(tag, bits, nq, nd, thr_v, K, workers=-1, probe=None):
    L.cmh_tc_set_workers(workers)
    db = engine.synth_codes(300 + bits, 0, nd, bits, dev)
    q = engine.synth_codes(400 + bits, 0, nq, bits, dev)
    tb = engine.TcBuffers(nq, [nd], bits, 1 << 22, dev)
    thr = torch.full((nq,), thr_v, dtype=torch.int32, device=dev)
    _cabi.check(L.cmh_tc_collect(engine._ptr(q.sign), nq, engine._ptr(db.sign), nd, bits, 0, engine._ptr(thr), K, 0, tb.seg_total,
                                 tb.seg_cap, engine._ptr(tb.cand), engine._ptr(tb.cnt), engine._ptr(tb.aux), engine._stream(dev)), "collect")
    torch.cuda.synchronize()
    L.cmh_tc_set_workers(-1)
    cnt = tb.cnt.cpu().numpy().astype(np.int64); cand = tb.cand.cpu().numpy().view(np.uint64)
    over = int((cnt > tb.seg_cap).sum())
    qs = q.sign.cpu().numpy().view(np.uint64)[:, 0]; ds = db.sign.cpu().numpy().view(np.uint64)[:, 0]
    n_missed = n_extra = n_total = 0; ex = []
    for qi in range(nq):
        dist = np.bitwise_count(qs[qi] ^ ds)
        want = set(np.nonzero(dist <= thr_v)[0].tolist()); got = {}
        for s in range(tb.seg_total):
            keys = cand[qi, s, :min(cnt[s, qi], tb.seg_cap)]
            for k in keys.tolist(): got[k & 0xffffffff] = k >> 33
        gs = set(got)
        n_total += len(want); n_missed += len(want - gs); n_extra += len(gs - want)
        if len(ex) < 6:
            for r in sorted(want - gs)[:2]: ex.append(("miss", qi, r, r % 256, int(dist[r])))
            for r in sorted(gs - want)[:2]: ex.append(("extra", qi, r, r % 256, int(dist[r]), "stored dist", got[r]))
    print(f"{tag}: bits={bits} thr={thr_v} K={K} workers={workers}: wanted {n_total}, missed {n_missed}, extra {n_extra}, overflowed segs {over}; {ex}", flush=True)
nq, nd = 256, 256 * 600
run("A 32-bit non-WK dense", 32, nq, nd, 8, 0)
run("B 32-bit WK forced, K=0", 32, nq, nd, 8, 0, workers=1)
run("C 32-bit non-WK forced, K>0", 32, nq, nd, 8, 1 << 30, workers=0)
run("D 32-bit non-WK sparse", 32, nq, nd, 4, 0)
run("E 64-bit non-WK very dense", 64, nq, nd, 24, 0)
run("F 64-bit non-WK forced K>0 very dense", 64, nq, nd, 24, 1 << 30, workers=0)
run("G 48-bit (64 wide) non-WK dense", 48, nq, nd, 14, 0)
