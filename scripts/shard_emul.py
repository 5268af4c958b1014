"""One shard of an N-GPU search on ONE GPU (loopback transport: cmh_comm_create_loopback): the per-shard device time of
every phase of cmh_topk_tc at world = N - all launches, the small kernels between them, the merge of N lists - without
the wire time of the collectives.  WORLD=8 python scripts/shard_emul.py"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import _cabi, engine, sharded

WORLD, RANK = int(os.environ.get("WORLD", 8)), int(os.environ.get("RANK", 0))
Q, D, K = int(os.environ.get("Q", 8192)), int(os.environ.get("D", 100_000_000)), 1000
dev = torch.device("cuda", 0)
L = _cabi.lib()


class Loop:
    def __init__(self, world, rank):
        self.world, self.rank = world, rank
        self.ptr = ctypes.POINTER(_cabi.Comm)()
        _cabi.check(L.cmh_comm_create_loopback(world, rank, ctypes.byref(self.ptr)), "cmh_comm_create_loopback")

    def handle(self):
        return self.ptr


comm = Loop(WORLD, RANK)
ranges, stripes = sharded.lockstep_stripes(D, WORLD, RANK)
rows = torch.cat([engine.synth_codes(4000, a, b - a, 64, dev).sign for a, b in ranges])
db = engine.PackedSet(rows, None, None, rows.shape[0], 64)
share = max(4096, 65536 * db.n // D)
smp_rows = db.sign[::max(1, db.n // share)].contiguous()
smp = engine.PackedSet(smp_rows, None, None, smp_rows.shape[0], 64)
qs = [engine.synth_codes(4001, i * Q, Q, 64, dev) for i in range(8)]
buffers = {}
fb = lambda sub: torch.full((sub.n, K), -1, dtype=torch.int64, device=dev)
for gather in (False, True):
    for i in range(3):
        engine.topk_tc(qs[i], db, K, ranges[0][0], sample=smp, comm=comm, nd_total=D, stripes=stripes, buffers=buffers,
                       gather=gather, exact_fallback=fb)
    torch.cuda.synchronize()
    st = {"time_phases": True}
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(3, 8):
        engine.topk_tc(qs[i], db, K, ranges[0][0], sample=smp, comm=comm, nd_total=D, stripes=stripes, buffers=buffers,
                       gather=gather, stats=st, exact_fallback=fb)
    b.record(); torch.cuda.synchronize()
    n = st["timed_searches"]
    print(f"world {WORLD} gather {gather}: {a.elapsed_time(b) / 5:.3f} ms/step; phases",
          {k: round(v / n, 3) for k, v in st["phase_ms_sum"].items()}, "collect", round(st["collect_ms_sum"] / n, 3),
          "launches", [round(x, 3) for x in st["launch_ms"]], "n_fail", st["n_fail"], "W", st["exch_width"],
          "cand/query", round(float(st["candidates"].float().mean()), 1))
