import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine
dev = torch.device("cuda", 0)
Q, K = 1024, 1000
for bits, D in ((16, 50_000_000), (16, 100_000_000), (32, 100_000_000)):
    db = engine.synth_codes(7000 + bits, 0, D, bits, dev)
    q = engine.synth_codes(7001 + bits, 0, Q, bits, dev)
    smp = engine.PackedSet(db.sign[::max(1, D // 65536)].contiguous(), None, None, -(-D // max(1, D // 65536)), bits)
    st, buf = {}, {}
    keys = engine.topk_tc(q, db, K, 0, sample=smp, stats=st, buffers=buf)
    sp = list(buf.values())[0]
    cnt = sp.cnt
    p = sp.plan
    print(bits, D, "n_fail", st["n_fail"], "thr", st["thr"][:6].tolist(), "thr_final", st["thr_final"][:6].tolist(),
          "cand/query", int(st["candidates"].float().mean()), "max seg cnt", int(cnt.max()), "seg_cap", p.seg_cap, "seg_total", p.seg_total,
          "spans", [(int(p.span_lo[i]), int(p.span_hi[i]), int(p.span_n_segs[i])) for i in range(p.n_spans)],
          "per-span max cnt", [int(cnt[int(p.span_seg_base[i]):int(p.span_seg_base[i]) + int(p.span_n_segs[i])].max()) for i in range(p.n_spans)],
          "kth dist2", int(keys[0, K - 1] >> 32), flush=True)
    del db, q, smp, buf, sp, cnt; torch.cuda.empty_cache()
