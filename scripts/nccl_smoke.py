"""Step-by-step check of the library's NCCL transport under torchrun (prints after every step, hard timeouts)."""
import ctypes, faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(90, exit=True)
import torch, torch.distributed as dist
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def say(*a):
    print(f"[{rank}] {time.strftime('%H:%M:%S')}", *a, flush=True)
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev); say("torch pg up")
from cmh_b200 import _cabi, engine, sharded
L = _cabi.lib()
comm = sharded.native_comm(None); say("native comm created", comm.contents.rank, comm.contents.world)
c = comm.contents
buf = torch.arange(8, dtype=torch.int32, device=dev) + 100 * rank
st = torch.cuda.Stream(dev)
with torch.cuda.stream(st):
    rc = c.all_reduce_u32(c.ctx, buf.data_ptr(), 8, 0, ctypes.c_void_p(st.cuda_stream)); st.synchronize()
say("all_reduce on a side stream rc", rc, buf.tolist())
rc = c.all_reduce_u32(c.ctx, buf.data_ptr(), 8, 0, None); torch.cuda.synchronize()
say("all_reduce on the NULL stream rc", rc, buf.tolist())
send = torch.full((world, 4), rank, dtype=torch.int64, device=dev); recv = torch.empty_like(send)
rc = c.all_to_all(c.ctx, send.data_ptr(), recv.data_ptr(), 32, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
say("all_to_all rc", rc, recv[:, 0].tolist())
g = torch.empty((world, 4), dtype=torch.int64, device=dev)
rc = c.all_gather(c.ctx, send[0].data_ptr(), g.data_ptr(), 32, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)); torch.cuda.synchronize()
say("all_gather rc", rc, g[:, 0].tolist())
# a small sharded search through cmh_topk_tc
D, Q, K = 4_000_000, 512, 100
ranges, stripes = sharded.lockstep_stripes(D, world, rank)
rows = torch.cat([engine.synth_codes(4000, a, b - a, 64, dev).sign for a, b in ranges])
db = engine.PackedSet(rows, None, None, rows.shape[0], 64)
from cmh_b200.index import HammingIndex
idx = HammingIndex(db, ranges[0][0], nd_total=D, stripes=stripes, assume_binary=True); say("index built")
q = engine.synth_codes(4001, 0, Q, 64, dev)
stt = {}
keys = idx.search_packed(q, K, stats=stt, gather=True); say("search done n_fail", stt["n_fail"], "W", stt["exch_width"])
whole = engine.synth_codes(4000, 0, D, 64, dev)
want = engine.topk_exact(q, whole, K, 0); say("equal to the exact single-GPU ranking:", bool(torch.equal(keys, want)))
sl = idx.search_packed(q, K, gather=False); lo, n = idx.query_slice(Q); say("slice equal:", bool(torch.equal(sl[:n], want[lo:lo + n])))
dist.barrier(); say("done"); dist.destroy_process_group()
