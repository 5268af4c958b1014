import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine, calc_utils as cu
from cmh_b200.index import HammingIndex
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
Q, D, K = 8192, 100_000_000, 1000
db = engine.synth_codes(4000, 0, D, 64, dev)
q_host = bench._host_float_codes(4001, 0, Q).pin_memory()
db_host = torch.empty((D, 1), dtype=torch.int64).pin_memory(); db_host.copy_(db.sign)
db_dev = torch.empty_like(db.sign)
keys_host = torch.empty((Q, K), dtype=torch.int64).pin_memory()
def timeit(fn, n=4):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return round((time.perf_counter() - t0) / n * 1e3, 2)
def copy_only():
    db_dev.copy_(db_host, non_blocking=True); torch.cuda.current_stream().synchronize()
def search_only():
    idx = HammingIndex.from_packed(db_dev, 64, 0, nd_total=D)
    k = idx.search_packed(cu.pack_codes(q_host.to(dev, non_blocking=True), dev), K)
    keys_host.copy_(k, non_blocking=True); torch.cuda.current_stream().synchronize()
def serial():
    db_dev.copy_(db_host, non_blocking=True); search_only()
def streamed(pieces):
    def f():
        idx = HammingIndex.from_packed_host(db_host, 64, 0, nd_total=D, out=db_dev, pieces=pieces)
        k = idx.search_packed(cu.pack_codes(q_host.to(dev, non_blocking=True), dev), K)
        keys_host.copy_(k, non_blocking=True); torch.cuda.current_stream().synchronize()
    return f
def build_only():
    idx = HammingIndex.from_packed_host(db_host, 64, 0, nd_total=D, out=db_dev, pieces=3); torch.cuda.synchronize()
print("copy_only", timeit(copy_only), "GB/s", round(0.8 / (timeit(copy_only) * 1e-3), 1))
print("search_only (resident db + q H2D + keys D2H)", timeit(search_only))
print("serial", timeit(serial))
for p in (1, 2, 3, 6): print("streamed pieces", p, timeit(streamed(p)))
print("build_only(3)", timeit(build_only))
t0 = time.perf_counter(); s = db_host[::1525].contiguous(); print("host sample gather ms", round((time.perf_counter() - t0) * 1e3, 2), s.shape)
