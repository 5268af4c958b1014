"""calc_map_k_matrix at the C2 shape: wall time per call and its kernels (run under ncu for the launch list)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import calc_utils as cu
from cmh_b200.synth import CONFIGS, make_case
name = os.environ.get("CFG", "c2-64")
shape = CONFIGS[name]
t = make_case(shape, clustered=True, zero_query_frac=0.01)
dev = torch.device("cuda", 0)
qB, rB = torch.from_numpy(t["q_img"]).to(dev), torch.from_numpy(t["r_txt"]).to(dev)
qLh, rLh = torch.from_numpy(t["q_lab"]), torch.from_numpy(t["r_lab"])
qL, rL = qLh.to(dev), rLh.to(dev)
for labels, tag in (((qL, rL), "device labels"), ((qLh, rLh), "host labels")):
    for _ in range(3): cu.clear_cache(); m = cu.calc_map_k_matrix(qB, rB, labels[0], labels[1], shape.k, 0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): cu.clear_cache(); m = cu.calc_map_k_matrix(qB, rB, labels[0], labels[1], shape.k, 0)
    dt = (time.perf_counter() - t0) / 10
    print(name, tag, "cold cache ms/call", round(dt * 1e3, 3), "map", float(m))
    for _ in range(10): m = cu.calc_map_k_matrix(qB, rB, labels[0], labels[1], shape.k, 0)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): m = cu.calc_map_k_matrix(qB, rB, labels[0], labels[1], shape.k, 0)
    dt = (time.perf_counter() - t0) / 10
    print(name, tag, "warm cache ms/call", round(dt * 1e3, 3))
