"""Per-rank GPU time of a sharded search, measured on ONE GPU: this process plays rank R of WORLD with a comm whose
collectives pretend the other shards hold statistically identical rows (sums scale by WORLD, gathers replicate).  The
results are meaningless (and fail verification); the kernel work per rank is the real thing, the NCCL latencies are not
in it."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine

WORLD = int(os.environ.get("WORLD", 8)); RANK = int(os.environ.get("RANK_EMUL", WORLD // 2))
Q, D, K = 8192, 100_000_000, 1000
dev = torch.device("cuda", 0)

class FakeComm:
    world, rank = WORLD, RANK
    def all_reduce_sum(self, t): return t * WORLD
    def all_reduce_max(self, t): return t
    def all_gather_stack(self, t): return t.unsqueeze(0).expand(WORLD, *t.shape).contiguous()
    def all_to_all(self, t): return t.clone()

n = D // WORLD
db = engine.synth_codes(4000, RANK * n, n, 64, dev); q = engine.synth_codes(4001, 0, Q, 64, dev)
share = max(4096, 65536 * n // D)
smp = engine.PackedSet(db.sign[::max(1, n // share)].contiguous(), None, None, 0, 64); smp.n = smp.sign.shape[0]
comm = FakeComm()

def run(label, n_it=5, **kw):
    bufs = {}
    for _ in range(3):
        engine.topk_tc(q, db, K, RANK * n, sample=smp, comm=comm, nd_total=D, buffers=bufs, exact_fallback=lambda sub: torch.zeros((sub.n, K), dtype=torch.int64, device=dev), **kw)
    stats = {"time_collect": True, "time_phases": True}
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n_it):
        engine.topk_tc(q, db, K, RANK * n, sample=smp, comm=comm, nd_total=D, buffers=bufs, stats=stats, exact_fallback=lambda sub: torch.zeros((sub.n, K), dtype=torch.int64, device=dev), **kw)
    e1.record(); torch.cuda.synchronize()
    pe = stats["phase_events"]; ph = {}
    for (n0, a), (n1, b) in zip(pe, pe[1:]):
        if n1 != "start": ph[n1] = ph.get(n1, 0.0) + a.elapsed_time(b) / n_it
    ev = stats["collect_events"]; per = len(ev) // 2 // n_it
    col = [sum(ev[2 * (i * per + j)].elapsed_time(ev[2 * (i * per + j) + 1]) for i in range(n_it)) / n_it for j in range(per)]
    print(f"{label:24s} {e0.elapsed_time(e1) / n_it:.3f} ms/step  " + " ".join(f"{k[:-5]}={v:.2f}" for k, v in ph.items()) +
          "  collect " + "+".join(f"{c:.2f}" for c in col) + f" = {sum(col):.2f}  cand/q {float(stats['candidates'].float().mean()):.0f}", flush=True)

run("contiguous shard", n_it=8)
ranges, stripes = __import__("cmh_b200.sharded", fromlist=["x"]).lockstep_stripes(D, WORLD, RANK)
rows = torch.cat([engine.synth_codes(4000, a, b - a, 64, dev).sign for a, b in ranges])
db = engine.PackedSet(rows, None, None, rows.shape[0], 64)
run("lockstep stripes", n_it=8, stripes=stripes)
