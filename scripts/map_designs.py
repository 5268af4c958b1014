"""The two counting passes per kernel design (0 tile, 1 generic warp, 2 lane) at the BASELINE.json shapes: device time of
pass 1 (cmh_eval_hist) and pass 2 (cmh_eval_rank) by CUDA events, and the mAP each design returns (must be identical)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cmh_b200 import calc_utils as cu, engine
from cmh_b200.synth import CONFIGS, make_case

dev = torch.device("cuda", 0)
names = os.environ.get("CFGS", "c1,c2-16,c2-32,c2-64,c3,c5").split(",")
designs = [int(x) for x in os.environ.get("DESIGNS", "0,2").split(",")]
out = {}
for name in names:
    shape = CONFIGS[name]
    t = make_case(shape, clustered=True, zero_query_frac=0.01)
    q, d = cu._prepare(torch.from_numpy(t["q_img"]).to(dev), torch.from_numpy(t["r_txt"]).to(dev),
                       torch.from_numpy(t["q_lab"]).to(dev), torch.from_numpy(t["r_lab"]).to(dev), 0)
    for design in designs:
        try:
            rp = engine.RankPass(q, d, need_labels=True, max_topn=0, design=design)
        except ValueError as e:
            out[f"{name}/d{design}"] = {"error": str(e)[:80]}
            continue
        th, tr = [], []
        for i in range(6):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(); rp.hist(); e[1].record(); ap_sum, n_rel, _ = rp.rank(shape.k); e[2].record()
            torch.cuda.synchronize()
            th.append(e[0].elapsed_time(e[1])); tr.append(e[1].elapsed_time(e[2]))
        ap, m = engine.finalize_map(ap_sum, n_rel, shape.k)
        out[f"{name}/d{design}"] = {"hist_ms": round(min(th[1:]), 4), "rank_ms": round(min(tr[1:]), 4), "map": float(m.cpu()[0]),
                                    "chunks": rp.plan.n_chunks, "chunk_rows": rp.plan.chunk_rows, "q_tile": rp.plan.q_tile}
        print(name, design, out[f"{name}/d{design}"], flush=True)
print(json.dumps(out))
