"""Time the pieces of the tensor-core top-K path at the bench scale (run on the GPU box)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import _cabi, engine

Q = int(os.environ.get("Q", 8192)); D = int(os.environ.get("D", 100_000_000)); K = 1000
STRIDE = int(os.environ.get("STRIDE", 384)); CAP = int(os.environ.get("CAP", 8192))
dev = torch.device("cuda", 0)
L = _cabi.lib()
db = engine.synth_codes(4000, 0, D, 64, dev)
q = engine.synth_codes(4001, 0, Q, 64, dev)
sample = engine.PackedSet(db.sign[::STRIDE].contiguous(), None, None, (D + STRIDE - 1) // STRIDE, 64)
st = engine._stream(dev); p = engine._ptr

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

rp = engine.RankPass(q, sample, need_labels=False)
t_hist = timed(lambda: rp.hist())
h_all, _ = rp.hist()
thr = torch.empty(Q, dtype=torch.int32, device=dev)
engine.check(L.cmh_topk_threshold(p(h_all), Q, 65, sample.n, D, K, p(thr), st))
cand = torch.empty((Q, CAP), dtype=torch.int64, device=dev); cnt = torch.empty(Q, dtype=torch.int32, device=dev)
keys = torch.empty((Q, K), dtype=torch.int64, device=dev)
ff = torch.empty(Q, dtype=torch.int32, device=dev); fc = torch.zeros(1, dtype=torch.int32, device=dev)
aux = torch.empty((Q, 8), dtype=torch.int32, device=dev)
t_static = timed(lambda: engine.check(L.cmh_tc_collect(p(q.sign), Q, p(db.sign), D, 64, 0, p(thr), 0, CAP, p(cand), p(cnt), p(aux), st)))
c_static = cnt.cpu().float()
t_collect = timed(lambda: engine.check(L.cmh_tc_collect(p(q.sign), Q, p(db.sign), D, 64, 0, p(thr), K, CAP, p(cand), p(cnt), p(aux), st)))
c = cnt.cpu().float()
t_final = timed(lambda: engine.check(L.cmh_topk_finalize(p(cand), p(cnt), p(aux), Q, CAP, K, D, p(keys), p(ff), p(fc), st)))
n_fail = int(fc.item())
# collect with an impossible threshold = pure GEMM + filter cost, no hits
thr0 = torch.full((Q,), -1, dtype=torch.int32, device=dev)
t_nohit = timed(lambda: engine.check(L.cmh_tc_collect(p(q.sign), Q, p(db.sign), D, 64, 0, p(thr0), 0, CAP, p(cand), p(cnt), p(aux), st)))
t_total = timed(lambda: engine.topk_tc(q, db, K, 0, sample=sample, cap=CAP))
import subprocess
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(json.dumps({"clocks_after": clk, "Q": Q, "D": D, "sample_rows": sample.n, "hist_sample_ms": t_hist, "collect_ms": t_collect, "collect_static_thr_ms": t_static, "cand_static_mean": float(c_static.mean()), "cand_static_max": float(c_static.max()),
                  "collect_nohit_ms": t_nohit, "finalize_ms": t_final, "topk_tc_total_ms": t_total,
                  "pairs_per_s_collect": Q * D / t_collect * 1e3, "pairs_per_s_total": Q * D / t_total * 1e3,
                  "cand_mean": float(c.mean()), "cand_max": float(c.max()), "cand_min": float(c.min()),
                  "n_fail": n_fail, "thr_min": int(thr.min()), "thr_max": int(thr.max())}))
