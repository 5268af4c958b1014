"""NS1 evidence: the two Hamming variants - XOR + POPC (`cmh_topk`, integer pipe) and +-1 int8 `tcgen05.mma`
(`cmh_topk_tc`, tensor pipe) - timed on the same top-K problem at every code length the reference's trainers emit
(16 / 32 / 64 / 128 bits, main.py:40, plus an odd one).  Keys are compared bit for bit; ms per search is the median of
REPS searches after warm-up, CUDA events on the launching stream.  JSON line at the end."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from cmh_b200 import engine
from cmh_b200.index import HammingIndex
Q, D, K = int(os.environ.get("Q", 2048)), int(os.environ.get("D", 20_000_000)), 1000
REPS = int(os.environ.get("REPS", 5))
dev = torch.device("cuda", 0)
out = {}
for bits in (16, 32, 48, 64, 96, 128):
    db = engine.synth_codes(7000 + bits, 0, D, bits, dev)
    qs = [engine.synth_codes(7001 + bits, i * Q, Q, bits, dev) for i in range(REPS + 2)]
    idx = HammingIndex(db, 0, nd_total=D, assume_binary=True)
    def timed(fn):
        ms = []
        for i in range(REPS + 2):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); r = fn(qs[i]); b.record(); torch.cuda.synchronize()
            if i >= 2: ms.append(a.elapsed_time(b))
        return float(np.median(ms)), r
    st = {}
    tc_ms, tc_keys = timed(lambda q: idx.search_packed(q, K, stats=st))
    popc_ms, popc_keys = timed(lambda q: engine.topk_exact(q, db, K))
    equal = bool(torch.equal(tc_keys, popc_keys))
    out[str(bits)] = {"popc_ms": round(popc_ms, 3), "tc_ms": round(tc_ms, 3), "ratio": round(popc_ms / tc_ms, 2), "keys_equal": equal,
                      "n_fail": int(st.get("n_fail", -1)), "popc_compares_per_s": Q * D / popc_ms * 1e3, "tc_compares_per_s": Q * D / tc_ms * 1e3}
    print(bits, out[str(bits)], flush=True)
    del db, idx, qs
    torch.cuda.empty_cache()
print(json.dumps({"variants_by_bits": {"queries": Q, "db_rows": D, "K": K, "by_bits": out}}))
