#!/usr/bin/env python
"""Benchmark of the retrieval-evaluation hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (host cores), same metric

Workload (`config.workload`): BASELINE.json config 4 - 64-bit codes, exact stable top-1000 retrieval against a
100,000,000-row database (the config the north-star targets are quoted on).  One *step* ranks one chunk of
`--queries` queries (default 8192, a slice of the config's 1M) against the whole database.  The database is
sharded by contiguous row ranges over the N GPUs (strong scaling: total work per step is fixed); every GPU runs the
tensor-core search (int8 tcgen05 GEMM + fused candidate filter) on its shard with global per-query thresholds
(all-reduced sample / pilot histograms), and the per-shard results are merged after one NCCL all-gather.

Metric: Hamming compares/s = queries x database rows / step time, whole job.
  value  device time (CUDA events), inputs already packed and resident in HBM
  e2e    through the public API with HOST buffers: pinned-host float query codes and the pinned-host packed
         database shard are copied H2D inside the timed region, queries are packed, ranked, merged and the keys
         are read back D2H
Also reported: roofline of the dominant kernel (tc_collect_kernel) against the int8 tensor peak, with the in-situ
ceilings of the same launch (MMA only, drain only, no hits); the reference's CPU path timed on the host cores
(`cpu_baseline`); mAP@ALL queries/s at the NUS-WIDE shape (config 2) as a secondary figure (`also`).
Synthetic data: counter-based uniform random codes (`cmh_synth_codes` / `synth.splitmix_rows`), seeded labels.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

BITS, TOPK, DB_ROWS, SEED = 64, 1000, 100_000_000, 4000
METRIC, UNIT = "hamming_compares_per_s", "compares/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["native", "reference"], default="native")
    ap.add_argument("--queries", type=int, default=8192, help="queries per step (chunk of the config's 1M)")
    ap.add_argument("--db-rows", type=int, default=DB_ROWS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    return ap.parse_args()


def config_dict(args, world):
    return {"workload": "c4: 64-bit exact stable top-1000 Hamming retrieval, 100M-row database, one query chunk per step",
            "bits": BITS, "topk": TOPK, "db_rows": args.db_rows, "queries_per_step": args.queries,
            "queries": "a different chunk of the config's 1M queries every step (chunk i = query rows i*Q..(i+1)*Q)",
            "pipelining": "searches run one after the other on one stream (same scratch); the host enqueues chunk i+1 before it reads the verdict of chunk i, so the GPU does not wait for Python between chunks",
            "sharding": f"database rows over {world} GPUs in lockstep stripes (3 global stripes x {world} contiguous pieces, rank r holds piece r of each); queries replicated; all-reduced threshold / prefix-rule histograms, NCCL all-to-all by query slice + merge + verify; results stay sharded by query slice (rank r keeps slice r)"
                        if world > 1 else "single GPU holds the whole database",
            "l2": "per-step working set (packed shard + candidate segments, >1 GB) exceeds the 126 MB L2; no explicit flush",
            "seed": SEED}


# ---------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi, during the timed region)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region through NVML (the library nvidia-smi
    itself uses; an in-process query every 50 ms instead of a polling nvidia-smi process, whose per-sample driver
    calls measurably slowed the kernels down).  Falls back to `nvidia-smi -lms` when pynvml is unavailable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []          # (arrival time, sm MHz, max MHz, power W, [reasons])
        self.proc = None
        self.stop_flag = threading.Event()
        self.how = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
            self.how = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.how = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.how = "nvidia-smi"
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read_smi, daemon=True)
        self.thread.start()

    def _poll_nvml(self):
        n = self.nvml
        try:
            mx = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:
            mx = float("nan")
        while not self.stop_flag.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                try:
                    mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.rows.append((time.perf_counter(), sm, mx, pw, [name for name, bit in self.REASONS if mask & bit]))
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def _read_smi(self):
        for line in self.proc.stdout:
            p = [x.strip() for x in line.split(",")]
            if len(p) >= 8:
                try:
                    self.rows.append((time.perf_counter(), float(p[1]), float(p[2]), float(p[3]),
                                      [name for (name, _), v in zip(self.REASONS, p[4:8]) if v.lower().startswith("active")]))
                except ValueError:
                    pass

    def window(self, t0: float, t1: float):
        """Only the samples that arrived inside [t0, t1] (the timed region) count."""
        self.t0, self.t1 = t0, t1

    def stop(self):
        self.stop_flag.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.how is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML and no nvidia-smi"]}
        t0, t1 = getattr(self, "t0", float("-inf")), getattr(self, "t1", float("inf"))
        inside = [r for r in self.rows if t0 <= r[0] <= t1 + 0.06]
        if not inside:                                   # a region shorter than the polling period: nearest samples
            inside = self.rows[-3:]
        sm = [r[1] for r in inside]
        reasons = sorted({x for r in inside for x in r[4]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max((r[2] for r in inside), default=None),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(inside), "how": self.how,
                "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own per-query path on the host cores
# ---------------------------------------------------------------------------------------------------------------
class NumaLocal:
    """Pinned host buffers of a rank should live on the NUMA node its GPU hangs off: pages of `cudaHostAlloc` are placed
    where the allocating thread runs (first touch), and an upload that crosses the socket interconnect is bounded by it -
    eight ranks streaming their shards at once notice.  Inside the block the process runs on the GPU's local CPUs; the
    previous affinity is restored on exit (the pages stay where they are).  Silent no-op when sysfs does not say."""

    def __init__(self, dev):
        self.info = {"node": None}
        self.cpus, self.prev = None, None
        try:
            pr = torch.cuda.get_device_properties(dev)
            bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
            self.info = {"gpu": bdf, "node": node}
            if node >= 0:
                cpus = set()
                for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
                allowed = os.sched_getaffinity(0)
                self.cpus = (cpus & allowed) or None
                self.info["local_cpus_allowed"] = len(cpus & allowed)
        except Exception as e:  # noqa: BLE001
            self.info["error"] = repr(e)[:80]

    def __enter__(self):
        if self.cpus:
            self.prev = os.sched_getaffinity(0)
            os.sched_setaffinity(0, self.cpus)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            os.sched_setaffinity(0, self.prev)
        return False


def _host_float_codes(seed, row0, n):
    from cmh_b200.synth import splitmix_rows
    w = splitmix_rows(seed, row0, n, 1, BITS)
    bits = np.unpackbits(w.view(np.uint8).reshape(n, 8), axis=1, bitorder="little")
    return torch.from_numpy(bits.astype(np.float32) * 2 - 1)


def reference_step(q: torch.Tensor, r: torch.Tensor, K: int) -> None:
    """What `calc_map_k_matrix` does per query to rank the database (utils/calc_utils.py:30-31): a float
    distance row and a full sort (forced stable), here truncated to the first K entries.  Runs the oracle's
    op-for-op restatement (`/root/reference` does not exist on the GPU box)."""
    from oracle import cmh_oracle as orc
    for i in range(q.shape[0]):
        dist = orc.hamming_dist(q[i], r).squeeze(0)
        order = torch.sort(dist, stable=True).indices
        _ = order[:K]


def run_reference_sample(n_queries: int, n_rows: int, steps: int, warmup: int):
    torch.set_num_threads(os.cpu_count() or 1)
    r = _host_float_codes(SEED, 0, n_rows)
    q = _host_float_codes(SEED + 1, 0, n_queries)
    for _ in range(warmup):
        reference_step(q[:1], r, TOPK)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_step(q, r, TOPK)
    dt = (time.perf_counter() - t0) / steps
    return n_queries * n_rows / dt, dt


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nq, nd = 8, 2_000_000
    value, dt = run_reference_sample(nq, nd, args.steps, max(1, args.warmup))
    cfg = config_dict(args, 1)
    cfg.update({"queries_timed_per_step": nq, "db_rows_timed": nd,
                "extrapolation": "the CPU arm ranks a bounded sample of the workload (8 queries x 2M rows per step); its "
                                 "compares/s is per pair - a full sort is O(D log D), so the per-pair cost at 100M rows is "
                                 "higher than what is measured here (the reported value flatters the CPU)",
                "sort": "stable (the contract); the shipped reference calls torch.sort with the default flag"})
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"per step {nq} queries x {nd} database rows of the same synthetic codes: float "
                                       "distance row + full stable sort per query (utils/calc_utils.py:30-31), torch CPU, "
                                       "all host threads; per-pair cost extrapolates linearly in queries"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------
def map_cpu_baseline(t, shape, n_queries):
    """The reference's own `calc_map_k_matrix` (utils/calc_utils.py:16-39; op-for-op oracle port, torch CPU, all host
    threads) on a query prefix: per-query cost is independent across queries, so queries/s extrapolates linearly."""
    from oracle import cmh_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    qB, rB = torch.from_numpy(t["q_img"][:n_queries]), torch.from_numpy(t["r_txt"])
    qL, rL = torch.from_numpy(t["q_lab"][:n_queries]), torch.from_numpy(t["r_lab"])
    out = {}
    for tag, stable in (("stable_sort", True), ("as_shipped_unstable_sort", False)):
        orc.ap_per_query_sorted(qB[:2], rB, qL[:2], rL, shape.k, stable=stable)
        t0 = time.perf_counter()
        orc.ap_per_query_sorted(qB, rB, qL, rL, shape.k, stable=stable)
        out[tag] = n_queries / (time.perf_counter() - t0)
    return {"value": out["stable_sort"], "unit": "queries/s", "as_shipped_value": out["as_shipped_unstable_sort"],
            "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first {n_queries} of {shape.n_query} queries against all {shape.n_db} rows (same synthetic codes and "
                      "labels), one repetition per sort flavour; per-query cost is independent across queries: linear "
                      "extrapolation to the whole query set"}


def bench_also(dev, local, hbm_peak, cpu=True):
    """Single-GPU secondary legs.  The reference's own calls (calc_map_k_matrix both directions, p_topK, pr_curve) at the
    shapes of configs 1-3 and the eval stage of config 5: device float codes + labels in, pack + two counting passes +
    host scalar out, pack cache cleared before every call.  Each with its roofline (the XOR+POPC compare against the
    integer-pipe peak measured live by `cmh_measure_popc_peak`) and the reference's CPU time beside it."""
    import ctypes
    from cmh_b200 import _cabi, calc_utils as cu, engine
    from cmh_b200.index import HammingIndex
    from cmh_b200.synth import CONFIGS, make_case
    lib = _cabi.lib()
    peak = ctypes.c_double(0.0)
    _cabi.check(lib.cmh_measure_popc_peak(4096, 5, ctypes.byref(peak), engine._stream(dev)), "cmh_measure_popc_peak")
    popc_peak = float(peak.value)

    def per_call_ms(fn, reps=5):
        for _ in range(2):
            cu.clear_cache(); out = fn()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            cu.clear_cache(); out = fn()
        torch.cuda.synchronize(dev)
        return (time.perf_counter() - t0) / reps * 1e3, out

    also = {"popc32_peak_per_s": popc_peak}
    cpu_queries = {"c1": 48, "c2-16": 16, "c2-32": 16, "c2-64": 24, "c3": 24, "c5": 24}
    for name in ("c1", "c2-16", "c2-32", "c2-64", "c3", "c5"):     # c5: the eval stage of config 5
        shape = CONFIGS[name]
        t = make_case(shape, clustered=True, zero_query_frac=0.01)
        qi, qt, ri, rt = (torch.from_numpy(t[k]).to(dev) for k in ("q_img", "q_txt", "r_img", "r_txt"))
        qL, rL = torch.from_numpy(t["q_lab"]).to(dev), torch.from_numpy(t["r_lab"]).to(dev)
        ms_i2t, m_i2t = per_call_ms(lambda: cu.calc_map_k_matrix(qi, rt, qL, rL, shape.k, local))
        ms_t2i, m_t2i = per_call_ms(lambda: cu.calc_map_k_matrix(qt, ri, qL, rL, shape.k, local))
        pairs = shape.n_query * shape.n_db
        popc = pairs * ((shape.bits + 31) // 32) * 2            # two counting passes recompute every distance
        entry = {"shape": f"{shape.n_query} x {shape.n_db}, {shape.bits}-bit, {shape.n_labels} labels, "
                          f"mAP@{'ALL' if shape.k is None else shape.k}",
                 "map_i2t": float(m_i2t), "map_t2i": float(m_t2i), "ms_per_call_i2t": ms_i2t, "ms_per_call_t2i": ms_t2i,
                 "map_queries_per_s": shape.n_query / (ms_i2t * 1e-3),
                 "compares_per_s": pairs / (ms_i2t * 1e-3),
                 "roofline": {"bound": "integer pipe (XOR+POPC)", "achieved": popc / (ms_i2t * 1e-3), "peak": popc_peak,
                              "unit": "POPC32/s", "frac": popc / (ms_i2t * 1e-3) / popc_peak if popc_peak else None,
                              "algorithmic": "ceil(bits/32) POPC32 per pair x 2 passes; whole call incl. pack, scan, "
                                             "finalize and the host read of the scalar"}}
        if shape.topn:
            ms_p, _ = per_call_ms(lambda: cu.p_topK(qi, rt, qL, rL, list(shape.topn), local))
            entry["p_topK_ms_per_call"] = ms_p
        if name == "c3":
            ms_pr, _ = per_call_ms(lambda: cu.pr_curve(qi, rt, qL, rL, local))
            entry["pr_curve_ms_per_call"] = ms_pr
        if cpu:
            entry["cpu_baseline"] = map_cpu_baseline(t, shape, cpu_queries[name])
        also[name] = entry
        del qi, qt, ri, rt, qL, rL

    # K1 at HBM scale: float32 codes [8M, 64] (2 GB) -> packed planes, float32 labels [8M, 24] -> masks
    n_rows = 8_000_000
    x = torch.randint(0, 2, (n_rows, 64), device=dev, dtype=torch.int8).float() * 2 - 1
    lab = (torch.rand((n_rows, 24), device=dev) < 0.15).float()
    cnt = torch.zeros(2, dtype=torch.int64, device=dev)
    neg = torch.zeros(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    pack = {}
    for pname, fn, rd, wr in (("pack_codes", lambda: engine.pack_codes_device(x, cnt), n_rows * 64 * 4, 2 * n_rows * 8),
                              ("pack_labels", lambda: engine.pack_labels_device(lab, neg), n_rows * 24 * 4, n_rows * 8)):
        ts = []
        for _ in range(6):
            flush.zero_()                                        # L2 flush between launches
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ms = min(ts[2:])
        pack[pname] = {"rows": n_rows, "ms": ms, "algorithmic_bytes": rd + wr,
                       "roofline": {"bound": "hbm", "achieved": (rd + wr) / ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": (rd + wr) / ms / 1e6 / hbm_peak}}
    also["pack"] = pack
    del x, lab, flush

    # the 128-bit tensor path (config 3's code length at retrieval scale): 8192 x 50M, top-1000, device-resident
    try:
        D128, Q128 = 50_000_000, 8192
        db = engine.synth_codes(SEED + 128, 0, D128, 128, dev)
        idx = HammingIndex(db, 0, group=False, nd_total=D128, assume_binary=True)
        qs = [engine.synth_codes(SEED + 129, i * Q128, Q128, 128, dev) for i in range(4)]
        st = {}
        idx.search_packed(qs[0], TOPK)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(1, 4):
            idx.search_packed(qs[i], TOPK, stats=st)
        b.record(); torch.cuda.synchronize(dev)
        ms = a.elapsed_time(b) / 3
        also["bits128"] = {"shape": f"{Q128} x {D128}, 128-bit, top-{TOPK}", "ms_per_step": ms, "n_fail": st.get("n_fail"),
                           "compares_per_s": Q128 * D128 / (ms * 1e-3),
                           "int8_tops": Q128 * D128 * 256 / (ms * 1e-3) / 1e12}
        del db, idx, qs
    except Exception as e:  # noqa: BLE001 - a secondary leg must not take the headline down
        also["bits128"] = {"error": repr(e)}
    try:
        also["c5_e2e"] = bench_c5(dev, local)
    except Exception as e:  # noqa: BLE001
        also["c5_e2e"] = {"error": repr(e)}
    return also


def bench_c5(dev, local):
    """BASELINE.json config 5 end to end: random-init CLIP ViT-B/32 + DCHMT hash head (bf16, no_grad, fused attention)
    encodes a 5,000-query and a 100,000-row retrieval set from HOST batches (pinned, double-buffered H2D), the head +
    argmax + pack + scatter is ONE kernel per modality and batch (cmh_hash_head_pack), then the reference's four
    calc_map_k calls (train/base.py:259-262) on the packed buffers.  Synthetic images / tokens / labels."""
    from cmh_b200 import calc_utils as cu
    from cmh_b200.synth import CONFIGS, make_labels
    from cmh_b200.valid_loop import DchmtModel, get_code_dchmt
    shape = CONFIGS["c5"]
    bits, batch, ctx = shape.bits, 500, 32
    torch.manual_seed(5005)
    model = DchmtModel(bits).to(dev).to(torch.bfloat16).eval()
    rng = np.random.default_rng(shape.seed)
    q_lab = torch.from_numpy(make_labels(rng, shape.n_query, shape.n_labels, shape.label_p, 0.01))
    r_lab = torch.from_numpy(make_labels(rng, shape.n_db, shape.n_labels, shape.label_p, 0.0))
    # a pool of distinct pinned host batches, cycled (synthetic data: the H2D traffic and the encoder work are real)
    pool = []
    g = torch.Generator().manual_seed(55)
    for _ in range(4):
        image = torch.randn(batch, 3, 224, 224, generator=g).to(torch.bfloat16).pin_memory()
        text = torch.randint(1, 49000, (batch, ctx), generator=g)
        text[:, 0] = 49406                                                     # start of text
        eot = torch.randint(8, ctx, (batch,), generator=g)
        text[torch.arange(batch), eot] = 49407                                 # end of text = the highest id
        for i in range(batch):
            text[i, int(eot[i]) + 1:] = 0
        pool.append((image, text.pin_memory()))

    def batches(n):
        for b0 in range(0, n, batch):
            m = min(batch, n - b0)
            image, text = pool[(b0 // batch) % len(pool)]
            yield image[:m], text[:m], torch.arange(b0, b0 + m).pin_memory()

    def run():
        t0 = time.perf_counter()
        q_img, q_txt = get_code_dchmt(model, batches(shape.n_query), shape.n_query, dev)
        r_img, r_txt = get_code_dchmt(model, batches(shape.n_db), shape.n_db, dev)
        torch.cuda.synchronize(dev)
        t1 = time.perf_counter()
        maps = [cu.calc_map_k_matrix(a, b, q_lab, r_lab, shape.k, local)
                for a, b in ((q_img, r_txt), (q_txt, r_img), (q_img, r_img), (q_txt, r_txt))]
        torch.cuda.synchronize(dev)
        t2 = time.perf_counter()
        return t1 - t0, t2 - t1, [float(m) for m in maps]

    get_code_dchmt(model, batches(2 * batch), 2 * batch, dev)                  # warm-up: kernels, cuBLAS plans
    torch.cuda.synchronize(dev)
    enc_s, eval_s, maps = run()
    n_items = shape.n_query + shape.n_db
    flops = n_items * (4.37e9 + 2.9e9) * 1.0                                   # ~ViT-B/32 image (50 tokens) + text (32 tokens) forward
    return {"shape": f"{shape.n_query} queries + {shape.n_db} database items, {bits}-bit DCHMT head, batches of {batch}",
            "encode_s": enc_s, "eval_ms_four_directions": eval_s * 1e3, "total_s": enc_s + eval_s,
            "items_per_s": n_items / enc_s, "approx_encoder_tflops": flops / enc_s / 1e12, "maps": maps,
            "h2d_bytes": n_items * (3 * 224 * 224 * 2 + ctx * 8 + 8),
            "note": "encoder: torch bf16 GEMMs + fused SDPA (library code) under no_grad; head + argmax + pack + scatter: "
                    "cmh_hash_head_pack; evaluation: packed buffers straight into calc_map_k_matrix (4 calls)"}


def bench_sharded_map(dev, rank, world, max_over_ranks, barrier):
    """N > 1: calc_map_k_matrix + precision@N + PR curve with the database sharded by contiguous row ranges
    (`cmh_map_k_sharded` over the library's NCCL transport) at a scaled NUS-WIDE shape, against the same call on ONE GPU
    (rank 0 ranks the whole database alone): per-query AP must agree within 1e-7."""
    from cmh_b200 import engine, sharded
    nq, nd, bits, nlab, topn = 2100, 20_000_000, 64, 21, (1, 100, 1000)

    def labels(seed, row0, n):          # ~12.5 % density per label, a pure function of the global row
        w = engine.synth_codes(seed, row0, n, 64, dev).sign
        for j in (1, 2):
            w = w & engine.synth_codes(seed + j, row0, n, 64, dev).sign
        return w & ((1 << nlab) - 1)

    q = engine.synth_codes(SEED + 21, 0, nq, bits, dev).with_labels(labels(SEED + 30, 0, nq), nlab)
    lo, hi = sharded.shard_bounds(nd, world, rank)
    shard = engine.synth_codes(SEED + 20, lo, hi - lo, bits, dev).with_labels(labels(SEED + 40, lo, hi - lo), nlab)
    comm = sharded.GroupComm(None)
    res = None
    for _ in range(2):
        res = sharded.map_k_sharded_native(q, shard, None, nd, topn, comm=comm, want_pr=True, ternary=False)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        res = sharded.map_k_sharded_native(q, shard, None, nd, topn, comm=comm, want_pr=True, ternary=False)
    b.record()
    barrier()
    ms = max_over_ranks(a.elapsed_time(b) / 3)
    out = {"shape": f"{nq} x {nd}, {bits}-bit, {nlab} labels, mAP@ALL + precision@{list(topn)} + PR curve",
           "n_gpus": world, "ms_per_call": ms, "map_queries_per_s": nq / (ms * 1e-3), "compares_per_s": nq * nd / (ms * 1e-3),
           "map": float(res["map"].cpu()[0])}
    ap_all = res["ap"].clone()
    del shard
    if rank == 0:
        whole = engine.synth_codes(SEED + 20, 0, nd, bits, dev).with_labels(labels(SEED + 40, 0, nd), nlab)
        one = None
        for _ in range(2):
            one = sharded.map_k_sharded_native(q, whole, None, nd, topn, comm=None, want_pr=True, ternary=False)
        a.record()
        one = sharded.map_k_sharded_native(q, whole, None, nd, topn, comm=None, want_pr=True, ternary=False)
        b.record(); torch.cuda.synchronize(dev)
        ms1 = a.elapsed_time(b)
        err = float((ap_all - one["ap"]).abs().max())
        out.update({"one_gpu_ms_per_call": ms1, "speedup_vs_one_gpu": ms1 / ms, "max_abs_ap_diff_vs_one_gpu": err,
                    "n_rel_equal": bool(torch.equal(res["n_rel"], one["n_rel"])),
                    "prec_max_abs_diff": float((res["prec"] - one["prec"]).abs().max()),
                    "pr_max_abs_diff": float(max((res["pr"][0] - one["pr"][0]).abs().max(), (res["pr"][1] - one["pr"][1]).abs().max()))})
        assert err < 1e-7 and out["n_rel_equal"], f"sharded mAP differs from the one-GPU result: {out}"
    barrier()
    return out


def main_native(args):
    import ctypes
    import torch.distributed as dist
    from cmh_b200 import _cabi, calc_utils as cu, engine, sharded
    from cmh_b200.index import HammingIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    Q, D, K = args.queries, args.db_rows, TOPK
    # One GPU holds the rows 0..D-1.  N GPUs hold them in LOCKSTEP STRIPES (`sharded.lockstep_stripes`): three global
    # stripes, each cut into N contiguous pieces, rank r holding piece r of every stripe - all shards walk the database
    # front to back together, so the exact prefix rule tightens as early as on one GPU.  (One contiguous range per
    # shard is supported too - `HammingIndex(db, index_base)` - but shard 0 then never benefits from the rule.)
    if world > 1:
        ranges, stripes = sharded.lockstep_stripes(D, world, rank)
        parts = [engine.synth_codes(SEED, a, b - a, BITS, dev) for a, b in ranges]
        rows = torch.cat([p.sign for p in parts])
        db = engine.PackedSet(rows, None, None, rows.shape[0], BITS)
        del parts
    else:
        ranges, stripes = [(0, D)], None
        db = engine.synth_codes(SEED, 0, D, BITS, dev)
    lo, hi = ranges[0][0], ranges[0][0] + db.n          # hi - lo = rows of this shard
    # a DIFFERENT chunk of the config's 1M queries every step (warm-up included): chunk i = query rows i*Q .. (i+1)*Q
    n_chunks = args.warmup + args.steps
    chunks = [engine.synth_codes(SEED + 1, i * Q, Q, BITS, dev) for i in range(n_chunks)]
    index = HammingIndex(db, lo, nd_total=D, stripes=stripes, assume_binary=True)

    # ---- device-resident timing ("value") ------------------------------------------------------------------
    # NVML sampling is started before the warm-up: its start-up takes driver locks and must not land in the timed
    # region; only the samples that arrive inside the region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()

    # Query chunks are searched one after the other on one stream.  The host stays one chunk ahead: chunk i+1 is enqueued
    # before the verdict of chunk i is read (`defer=True`), so the GPU does not idle while Python prepares the next call -
    # the searches themselves do not overlap (same stream, same scratch).  With N GPUs the result stays SHARDED BY QUERY
    # SLICE: rank r merges, verifies and keeps the keys of its slice of the chunk (`gather=False`; an all-gather of the
    # merged keys is one flag away and is what `search_packed` does by default).
    step_marks = []                              # an event behind every timed search: one-off stalls show in the line

    def run_steps(first, n, stats=None, marks=None):
        pending, keys = None, None
        for i in range(first, first + n):
            h = index.search_packed(chunks[i], K, stats=stats, gather=False, defer=True)
            if marks is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
            if pending is not None:
                keys = pending()
            pending = h
        return pending() if pending is not None else keys

    run_steps(0, args.warmup)
    barrier()
    t_region0 = time.perf_counter()
    launches0 = lib.cmh_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keys = run_steps(args.warmup, args.steps, marks=step_marks)
    e1.record()
    barrier()
    launches = lib.cmh_launch_count() - launches0
    if rank == 0:
        sampler.window(t_region0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    step_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    step_each = [round(a_.elapsed_time(b_), 3) for a_, b_ in zip([e0] + step_marks[:-1], step_marks)]
    value = Q * D / (step_ms * 1e-3)
    # per-phase device times (the library's own CUDA events, cmh_tc_timing) in a separate pass over the timed chunks, each
    # search resolved before the next: reading the events is a host sync that the timed region above does without
    stats = {"time_collect": True, "time_phases": True}
    n_fail_total = 0
    for i in range(args.warmup, args.warmup + min(args.steps, 5)):
        index.search_packed(chunks[i], K, stats=stats, gather=False)
        n_fail_total += stats.get("n_fail", 0)

    # ---- parity of what was just timed: a subsample of the LAST timed chunk re-ranked by the exact (popc, two-pass)
    # sharded path - every rank's slice contributes queries - must give the same keys bit for bit ------------------
    last = chunks[n_chunks - 1]
    per_rank = -(-Q // world)
    n_check = 64
    take = max(1, n_check // world)
    rows_check = torch.cat([torch.arange(r * per_rank, min(Q, r * per_rank + take)) for r in range(world)]).to(dev)
    sub = engine.PackedSet(last.sign.index_select(0, rows_check).contiguous(), None, None, int(rows_check.numel()), BITS)
    exact = sharded.topk_sharded(sub, db, K, lo, None, stripes=stripes)            # identical on every rank
    q_lo, q_n = index.query_slice(Q)
    mine = ((rows_check >= q_lo) & (rows_check < q_lo + q_n)).nonzero().squeeze(1)
    equal = bool(torch.equal(keys.index_select(0, rows_check[mine] - q_lo), exact.index_select(0, mine)))
    eq_t = torch.tensor([1 if equal else 0, int(mine.numel())], dtype=torch.int64, device=dev)
    if world > 1:
        flag = eq_t[:1].clone(); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        cnt = eq_t[1:].clone(); dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        eq_t = torch.cat([flag, cnt])
    parity = {"queries": int(eq_t[1]), "equal": bool(int(eq_t[0])), "n_fail": int(n_fail_total),
              "against": "exact two-pass popc path (cmh_topk per shard + all-gather + merge) on the same chunk"}
    if not parity["equal"]:
        raise AssertionError(f"tensor-core keys differ from the exact path on the timed chunk: {parity}")

    # ---- dominant kernel (tc_collect_kernel: pilot + main launches of every step), timed live by the events the
    # library records around its launches on the launching stream (cmh_tc_timing) ----------------------------------
    n_timed = max(1, stats.get("timed_searches", 1))
    phases = {k_: v / n_timed for k_, v in stats.get("phase_ms_sum", {}).items()}
    collect_ms = stats.get("collect_ms_sum", 0.0) / n_timed
    n_collect = stats.get("n_collect", 0)
    pairs_shard = Q * (hi - lo)
    ops_per_pair = 2 * BITS                                   # 2 * B_pad int8 ops per pair
    achieved_tops = pairs_shard * ops_per_pair / (collect_ms * 1e-3) / 1e12 if collect_ms else None
    # in-situ ceilings of the same launch with parts of the pipeline disabled (cmh_tc_probe)
    ceilings = {}
    tb = engine.TcBuffers(Q, [hi - lo], BITS, engine.TC_DEFAULT_CAP, dev)
    thr_never = torch.full((Q,), -1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)
    for name, mode in (("no_hits", 0), ("mma_only", 2), ("drain_only", 1)):
        ts = []
        for _ in range(3):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            _cabi.check(lib.cmh_tc_probe(engine._ptr(last.sign), Q, engine._ptr(db.sign), hi - lo, BITS,
                                         engine._ptr(thr_never), tb.seg_total, tb.seg_cap, engine._ptr(tb.cand),
                                         engine._ptr(tb.cnt), engine._ptr(tb.aux), mode, engine._stream(dev)), "cmh_tc_probe")
            b.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ceilings[name + "_ms"] = min(ts)
    del tb

    # ---- end to end through the public API with host buffers ------------------------------------------------
    # per step: this step's float32 query codes (pinned host) H2D + pack; the packed shard (pinned host) uploaded in row
    # ranges on a copy stream and scanned range by range as it lands; search; this rank's slice of the keys D2H
    numa = NumaLocal(dev)
    with numa:                                   # pinned pages on the GPU's NUMA node
        q_hosts = [_host_float_codes(SEED + 1, i * Q, Q).pin_memory() for i in range(n_chunks)]
        db_host = torch.empty((hi - lo, 1), dtype=torch.int64).pin_memory()
        db_host.copy_(db.sign)
    # Software pipeline over the steps, all through public calls: DEPTH + 1 device buffers for the shard, the upload of
    # step i + DEPTH (copy stream) is enqueued when step i is, so it runs under the scans in front of it; queries packed straight from
    # pinned host memory on a side stream (the kernel reads the float codes over the link: nothing queues behind the shard
    # on the copy engine, and the pack's counter read waits for 2 MB, not for the scan in front of it); keys D2H on a third
    # stream as soon as the step's verdict is in.  Every step still moves ITS queries and ITS copy of the shard H2D and
    # its keys D2H inside the timed region.
    DEPTH = int(os.environ.get("CMH_E2E_DEPTH", "1"))
    NB = DEPTH + 1
    PIECES = int(os.environ.get("CMH_E2E_PIECES", "0")) or None
    db_bufs = [torch.empty_like(db.sign) for _ in range(NB)]
    with numa:
        keys_hosts = [torch.empty((per_rank if world > 1 else Q, K), dtype=torch.int64).pin_memory() for _ in range(2)]
    main_stream = torch.cuda.current_stream(dev)
    q_stream, d2h_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    buf_free = [None] * NB                       # event behind the last search that read db_bufs[j]
    upload_ms, search_ev = [], []

    def e2e_upload(i):
        return HammingIndex.from_packed_host(db_host, BITS, lo, nd_total=D, out=db_bufs[i % NB], stripes=stripes,
                                             out_free=buf_free[i % NB], pieces=PIECES)

    def e2e_search(i, idx):
        with torch.cuda.stream(q_stream):        # pinned host codes: the pack kernel reads them straight over the link
            qp = cu.pack_codes(q_hosts[i], dev)
        qp.sign.record_stream(main_stream)
        main_stream.wait_stream(q_stream)
        e_a = torch.cuda.Event(enable_timing=True)
        e_a.record(main_stream)
        h = idx.search_packed(qp, K, gather=False, defer=True)
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(main_stream)
        buf_free[i % NB] = ev
        search_ev.append((e_a, ev))
        return h

    def e2e_resolve(i, h):
        k = h()                                  # verdict of step i (its scan, exchange and merge are complete)
        if getattr(h, "verdict", {}).get("n_fail", 1):
            d2h_stream.wait_stream(main_stream)  # redone queries were patched in on the main stream
        with torch.cuda.stream(d2h_stream):
            keys_hosts[i % 2].copy_(k, non_blocking=True)
        k.record_stream(d2h_stream)

    def e2e_run(first, n):
        last = first + n - 1
        idxs = {i: e2e_upload(i) for i in range(first, min(last, first + DEPTH - 1) + 1)}
        h_prev = None
        for i in range(first, last + 1):
            idx = idxs.pop(i)
            h = e2e_search(i, idx)
            if i + DEPTH <= last:
                # (after the search: the index's sample gather goes to the main stream and must not sit in front of it)
                idxs[i + DEPTH] = e2e_upload(i + DEPTH)
            if h_prev is not None:
                e2e_resolve(i - 1, h_prev)
            h_prev = h
            upload_ms.append(idx.upload_events)
        e2e_resolve(last, h_prev)
        torch.cuda.synchronize(dev)

    e2e_run(0, max(1, min(n_chunks, max(NB, args.warmup))))     # (never beyond the chunks that exist: --warmup 0 --steps 1)
    upload_ms.clear()
    search_ev.clear()
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.warmup, args.steps)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
    upload_avg = max_over_ranks(sum(a.elapsed_time(b) for a, b in upload_ms) / max(1, len(upload_ms)))
    # device time of the searches alone inside the e2e region (main-stream events around each): what is left of a step is
    # the pack, the waits for the link and the gaps between searches
    e2e_search_ms = max_over_ranks(sum(a.elapsed_time(b) for a, b in search_ev) / max(1, len(search_ev)))
    e2e_span_ms = max_over_ranks(search_ev[0][0].elapsed_time(search_ev[-1][1]) / max(1, len(search_ev)))
    # the library's own phase events for searches issued the e2e way (fresh index over an uploaded buffer; un-pipelined pass)
    st_e = {"time_collect": True, "time_phases": True}
    for i in range(args.warmup, args.warmup + min(args.steps, 3)):
        idx_e = e2e_upload(i)
        qp_e = cu.pack_codes(q_hosts[i], dev)
        idx_e.search_packed(qp_e, K, stats=st_e, gather=False)
    n_e = max(1, st_e.get("timed_searches", 1))
    e2e_phases = {k_: v / n_e for k_, v in st_e.get("phase_ms_sum", {}).items()}
    e2e_phases["launch_ms"] = st_e.get("launch_ms")
    keys_host = keys_hosts[(args.warmup + args.steps - 1) % 2]
    assert torch.equal(keys_host, keys.cpu()), "e2e keys differ from the device-resident run"
    h2d = q_hosts[0].numel() * 4 + db_host.numel() * 8
    d2h = keys_host.numel() * 8

    # ---- N > 1: the reference's own call, sharded: calc_map_k_matrix + precision@N + PR curve at a
    # scaled NUS-WIDE shape through cmh_map_k_sharded over the library's NCCL transport, against the one-GPU result ----
    sharded_map = None
    if world > 1 and not args.no_also:
        sharded_map = bench_sharded_map(dev, rank, world, max_over_ranks, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)

    # ---- secondary legs, single GPU: the reference's own evaluation calls at the shapes of configs 1-3 (+ the eval
    # stage of config 5), K1 at HBM scale, the 128-bit tensor path -----------------------------------------------
    also = None
    if not args.no_also and world == 1:          # (the other ranks have left: nothing collective may follow)
        also = bench_also(dev, local, hbm_peak, cpu=not args.no_cpu_baseline)

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        nq, nd = 8, 2_000_000
        v, dt = run_reference_sample(nq, nd, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{nq} queries x {nd} rows of the same synthetic codes, 2 timed repetitions: float distance row + "
                         "full stable sort per query (utils/calc_utils.py:30-31), torch CPU, all host threads"}

    # int8 tensor peak: tcgen05 kind::i8 retires twice the MACs per clock of kind::f16, so 2 x the measured dense bf16
    # rate (burst figure: the kernel is timed alone, launch by launch)
    bf16 = peaks.get("bf16_tflops")
    i8_peak = 2.0 * bf16 if bf16 else 2.0 * 1500.0
    traffic, traffic_capture = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic, traffic_capture = tj.get("tc_collect_kernel"), tj.get("tc_collect_kernel_capture")
    except Exception:
        pass
    algo_bytes = (hi - lo) * 8 + Q * 8                       # packed shard + packed queries, read once
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "ms_each_step": step_each, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "s8 (+-1 int8 tcgen05 MMA, int32 accumulate; exact integer distances and ranks)", "data": "synthetic",
        "config": config_dict(args, world),
        "e2e": {"value": Q * D / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "shard_upload_ms": upload_avg, "pipeline_depth": DEPTH, "pinned_host_numa": numa.info,
                "search_ms_per_step": e2e_search_ms, "search_phase_ms": e2e_phases, "first_search_start_to_last_search_end_ms_per_step": e2e_span_ms,
                "note": "per step and rank: pinned-host packed database shard (uploaded in row ranges on a copy stream into "
                        "one of two alternating device buffers, so the upload of step i+1 runs under the scan of step i) + "
                        "this step's float32 query codes H2D, index build (device-side sample of the first range), pack, "
                        "one cmh_topk_tc call (scan, all-to-all by query slice, merge + verify), this rank's slice of the "
                        "top-K keys D2H; steps are software-pipelined, all copies inside the timed region"},
        "gpu_launches": int(launches),
        "parity_check": parity,
        "phase_ms_per_step": phases,
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": "tc_collect_kernel (int8 tcgen05 GEMM + fused top-K candidate filter; "
                                         f"{n_collect:g} launches per step: pilot rows + the spans between the prefix-rule cuts)",
            "achieved": achieved_tops, "peak": i8_peak, "unit": "TOP/s (int8)",
            "frac": achieved_tops / i8_peak if achieved_tops else None,
            "peak_source": ("2 x MEASURED_PEAKS.json bf16_tflops (burst)" if bf16 else "fallback 2 x 1500 TFLOP/s") +
                           ": kind::i8 has twice the MAC rate of kind::f16",
            "algorithmic_ops_per_pair": ops_per_pair, "pairs_per_launch_set": pairs_shard,
            "kernel_ms_per_step": collect_ms, "phase_ms_per_step": phases,
            "in_situ_ceilings_ms": ceilings,
            "note": "the kernel issues 5 K-steps per 4 algorithmic ones (the per-query threshold rides in a bias K-step); "
                    "mma_only / drain_only / no_hits are the same launch over the whole shard with the TMEM drain / the "
                    "MMAs / the hit path disabled.  mma_only is the tensor pipe's own ceiling for this instruction "
                    "stream (its issued rate, 1.25 x algorithmic ops / mma_only time, is above the cuBLAS-bf16-derived "
                    "peak used here: kind::i8 M128 N128 K32 sustains close to the nominal 4.5 POP/s), so frac is "
                    "measured against the driver's denominator, not against what the pipe can do",
            "issued_tops_mma_only": (pairs_shard * ops_per_pair * 1.25 / (ceilings["mma_only_ms"] * 1e-3) / 1e12
                                     if ceilings.get("mma_only_ms") else None),
            "traffic": traffic, "traffic_capture": traffic_capture,   # ncu dram bytes of ONE captured launch of the set
            "hbm": {"achieved": algo_bytes / (collect_ms * 1e-3) / 1e9 if collect_ms else None, "peak": hbm_peak,
                    "unit": "GB/s", "peak_source": "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                    "algorithmic_bytes_per_launch_set": algo_bytes}},
        "cpu_baseline": cpu,
        "also": also,
        "sharded_map": sharded_map,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    # a hang must leave evidence: every 4 minutes without finishing, every thread's stack goes to stderr
    import faulthandler
    faulthandler.dump_traceback_later(240, repeat=True)
    args = parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_native(args)


if __name__ == "__main__":
    main()
