"""Time the pieces of the tensor-core top-K path at the bench scale (run on the GPU box)."""
import ctypes, sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import _cabi, engine

Q = int(os.environ.get("Q", 8192)); D = int(os.environ.get("D", 100_000_000)); K = 1000
BITS = int(os.environ.get("BITS", 64))
STRIDE = int(os.environ.get("STRIDE", 384)); CAP = int(os.environ.get("CAP", engine.TC_DEFAULT_CAP))
PROBES = os.environ.get("PROBES", "1") == "1"
dev = torch.device("cuda", 0)
L = _cabi.lib()
db = engine.synth_codes(4000, 0, D, BITS, dev)
q = engine.synth_codes(4001, 0, Q, BITS, dev)
sample = engine.PackedSet(db.sign[::STRIDE].contiguous(), None, None, (D + STRIDE - 1) // STRIDE, BITS)
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
b = engine.TcBuffers(Q, [D], BITS, CAP, dev)
engine.check(L.cmh_topk_threshold(p(h_all), Q, BITS + 1, sample.n, D, K, p(b.thr), st))
keys = torch.empty((Q, K), dtype=torch.int64, device=dev)


def collect(k, thr):
    engine.check(L.cmh_tc_collect(p(q.sign), Q, p(db.sign), D, BITS, 0, p(thr), k, 0, b.seg_total, b.seg_cap,
                                  p(b.cand), p(b.cnt), p(b.aux), st))


t_static = timed(lambda: collect(0, b.thr))
c_static = b.cnt.sum(0).cpu().float(); seg_static_max = int(b.cnt.max())
t_collect = timed(lambda: collect(K, b.thr))
c = b.cnt.sum(0).cpu().float(); seg_max = int(b.cnt.max())
t_final = timed(lambda: engine.check(L.cmh_topk_finalize(p(b.cand), p(b.cnt), p(b.aux), p(b.thr), Q, b.seg_total, b.seg_cap, K, D, 0,
                                                         p(keys), p(b.fail_flags), p(b.fail_count), st)))
n_fail = int(b.fail_count.item())
# exact thresholds (the K-th distance itself): the least work the hit path can be given
cum = torch.cumsum(torch.zeros(1), 0)
kth = (keys[:, K - 1] >> 33).to(torch.int32).contiguous()
t_exact_thr = timed(lambda: collect(0, kth))
c_exact = b.cnt.sum(0).cpu().float()
out = {"Q": Q, "D": D, "bits": BITS, "seg_total": b.seg_total, "seg_cap": b.seg_cap, "sample_rows": sample.n,
       "hist_sample_ms": t_hist, "collect_ms": t_collect, "collect_static_thr_ms": t_static,
       "collect_exact_thr_ms": t_exact_thr, "cand_exact_mean": float(c_exact.mean()),
       "cand_static_mean": float(c_static.mean()), "cand_static_max": float(c_static.max()),
       "seg_static_max": seg_static_max, "seg_max": seg_max, "finalize_ms": t_final,
       "pairs_per_s_collect": Q * D / t_collect * 1e3, "cand_mean": float(c.mean()), "cand_max": float(c.max()),
       "cand_min": float(c.min()), "n_fail": n_fail, "thr_min": int(b.thr.min()), "thr_max": int(b.thr.max())}
if PROBES:
    # in-situ ceilings: impossible threshold (no hits) with parts of the pipeline disabled
    thr0 = torch.full((Q,), -1, dtype=torch.int32, device=dev)
    for name, mode in (("nohit", 0), ("no_mma", 1), ("mma_only", 2), ("drain_noscan", 4), ("no_mma_noscan", 5),
                       ("no_mma_no_drain", 3)):
        t = timed(lambda: engine.check(L.cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(thr0), b.seg_total, b.seg_cap,
                                                      p(b.cand), p(b.cnt), p(b.aux), mode, st)))
        out[f"probe_{name}_ms"] = t
        out[f"probe_{name}_pairs_per_clk_sm"] = Q * D / (t * 1e-3) / 148 / 1.965e9
if os.environ.get("HITPROBES", "0") == "1":
    for name, mode in (("hit_full", 0), ("hit_nostore", 8), ("hit_nowork", 16), ("hit_nopark", 32)):
        t = timed(lambda: engine.check(L.cmh_tc_probe(p(q.sign), Q, p(db.sign), D, BITS, p(b.thr), b.seg_total, b.seg_cap,
                                                      p(b.cand), p(b.cnt), p(b.aux), mode, st)))
        out[f"probe_{name}_ms"] = t
stats = {}
t_total = timed(lambda: engine.topk_tc(q, db, K, 0, sample=sample, cap=CAP, stats=stats))
out["topk_tc_total_ms"] = t_total
out["pairs_per_s_total"] = Q * D / t_total * 1e3
out["total_cand_mean"] = float(stats["candidates"].float().mean()); out["total_n_fail"] = stats["n_fail"]
out["pilot_rows"] = stats["pilot_rows"]
t_nopilot = timed(lambda: engine.topk_tc(q, db, K, 0, sample=sample, cap=CAP, pilot=0))
out["topk_tc_nopilot_ms"] = t_nopilot
import subprocess
out["clocks_after"] = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                      "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
print(json.dumps(out))
