"""K1 (sign + bit-pack) at HBM scale: float32 codes [N, bits] -> packed planes, labels [N, L] -> masks; GB/s per launch
(CUDA events) against MEASURED_PEAKS.json hbm_gbs.  Run under ncu for dram bytes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import engine

dev = torch.device("cuda", 0)
N, BITS, L = int(os.environ.get("N", 8_000_000)), int(os.environ.get("BITS", 64)), int(os.environ.get("L", 24))
x = (torch.randint(0, 2, (N, BITS), device=dev, dtype=torch.int8).float() * 2 - 1)
lab = (torch.rand((N, L), device=dev) < 0.15).float()
cnt = torch.zeros(2, dtype=torch.int64, device=dev)
neg = torch.zeros(1, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for name, fn, rd, wr in (("pack_codes", lambda: engine.pack_codes_device(x, cnt), N * BITS * 4, 2 * N * ((BITS + 63) // 64) * 8),
                         ("pack_labels", lambda: engine.pack_labels_device(lab, neg), N * L * 4, N * ((L + 63) // 64) * 8)):
    ts = []
    for i in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts[3:])
    out[name] = {"rows": N, "ms": ms, "read_bytes": rd, "write_bytes": wr, "gbs": (rd + wr) / ms / 1e6}
print(json.dumps(out))
