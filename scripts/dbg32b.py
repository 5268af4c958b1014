"""Bare cmh_tc_collect launch against brute force: which (query, row) pairs at dist <= thr are missed?"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import engine, _cabi
from cmh_b200.synth import splitmix_rows
dev = torch.device("cuda", 0)
L = _cabi.lib()
for bits, nq, nd, thr_v, K in ((32, 512, 256 * 600, 8, 0), (32, 512, 256 * 600, 8, 1000000), (64, 512, 256 * 600, 20, 0), (16, 512, 256 * 600, 3, 0), (32, 100, 256 * 600 + 77, 8, 0)):
    db = engine.synth_codes(300 + bits, 0, nd, bits, dev)
    q = engine.synth_codes(400 + bits, 0, nq, bits, dev)
    tb = engine.TcBuffers(nq, [nd], bits, 1 << 20, dev)
    thr = torch.full((nq,), thr_v, dtype=torch.int32, device=dev)
    _cabi.check(L.cmh_tc_collect(engine._ptr(q.sign), nq, engine._ptr(db.sign), nd, bits, 0, engine._ptr(thr), K, 0, tb.seg_total,
                                 tb.seg_cap, engine._ptr(tb.cand), engine._ptr(tb.cnt), engine._ptr(tb.aux), engine._stream(dev)), "collect")
    torch.cuda.synchronize()
    cnt = tb.cnt.cpu().numpy().astype(np.int64)          # [segs][nq]
    cand = tb.cand.cpu().numpy().view(np.uint64)         # [nq][segs][cap]
    assert cnt.max() <= tb.seg_cap, (cnt.max(), tb.seg_cap)
    qs = q.sign.cpu().numpy().view(np.uint64)[:, 0]; ds = db.sign.cpu().numpy().view(np.uint64)[:, 0]
    n_missed = n_extra = n_total = 0
    missed = []
    for qi in range(nq):
        dist = np.bitwise_count(qs[qi] ^ ds)
        want = set(np.nonzero(dist <= thr_v)[0].tolist())
        got = set()
        for s in range(tb.seg_total):
            keys = cand[qi, s, :cnt[s, qi]]
            got.update((keys & np.uint64(0xffffffff)).astype(np.int64).tolist())
        n_total += len(want); n_missed += len(want - got); n_extra += len(got - want)
        for r in sorted(want - got)[:3]:
            missed.append((qi, r, r % 256, int(dist[r])))
    print(f"bits={bits} nq={nq} nd={nd} thr={thr_v} K={K}: wanted {n_total}, missed {n_missed}, extra {n_extra}; e.g. (query, row, row%256, dist) {missed[:12]}")
