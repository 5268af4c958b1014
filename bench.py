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
            "pipelining": "none: one query chunk at a time, resolved and verified before the next is enqueued",
            "sharding": f"database rows over {world} GPUs in lockstep stripes (3 global stripes x {world} contiguous pieces, rank r holds piece r of each); queries replicated; all-reduced threshold / prefix-rule histograms, NCCL all-to-all by query slice + merge + all-gather"
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((r[2] for r in inside), default=None),
                "power_w_max": max((r[3] for r in inside), default=None), "samples": len(inside), "how": self.how,
                "reasons": reasons}


# ---------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own per-query path on the host cores
# ---------------------------------------------------------------------------------------------------------------
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
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args, 1),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"per step {nq} queries x {nd} database rows of the same synthetic codes: float "
                                       "distance row + full stable sort per query (utils/calc_utils.py:30-31), torch CPU, "
                                       "all host threads; per-pair cost extrapolates linearly in queries"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------------------
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
    q_packed = engine.synth_codes(SEED + 1, 0, Q, BITS, dev)
    index = HammingIndex(db, lo, nd_total=D, stripes=stripes)

    # ---- device-resident timing ("value") ------------------------------------------------------------------
    # nvidia-smi is started before the warm-up: its NVML start-up takes driver locks and must not land in the timed
    # region; only the samples that arrive inside the region are reported
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # One query chunk at a time.  `search_packed_async` (two chunks in flight on alternating streams) is available, but
    # with one long-lived 768-thread CTA per SM the small kernels and the NCCL kernels between the launches of one chunk
    # only get SM slots at CTA boundaries of the other chunk's scan: measured -25 % at 2 GPUs, and a loss on a single GPU
    # too once the search became a chain of six to seven launches with refinement kernels in between.
    pipelined = False

    def run_steps(n, stats):
        pending, keys = None, None
        for _ in range(n):
            if not pipelined:
                keys = index.search_packed(q_packed, K, stats=stats)
                continue
            h = index.search_packed_async(q_packed, K, stats=stats)
            if pending is not None:
                keys = pending.result()
            pending = h
        return pending.result() if pending is not None else keys

    # warm-up on the timed code path (kernels loaded).  Its result is not kept: a third live [Q, K] key buffer next to
    # the two the loop alternates between sends the caching allocator to cudaMalloc in the middle of the timed region
    # (measured: 1-70 ms in the second timed step, wherever the warm-up ended).
    run_steps(args.warmup, {"time_collect": True, "time_phases": True})
    barrier()
    t_region0 = time.perf_counter()
    launches0 = lib.cmh_launch_count()
    stats = {"time_collect": True, "time_phases": True}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keys = run_steps(args.steps, stats)
    e1.record()
    barrier()
    launches = lib.cmh_launch_count() - launches0
    if rank == 0:
        sampler.window(t_region0, time.perf_counter())
    clocks = sampler.stop() if rank == 0 else None
    step_ms = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = Q * D / (step_ms * 1e-3)

    # ---- dominant kernel (tc_collect_kernel: pilot + main launch of every step), timed live by the events the
    # search recorded around its launches on the launching stream ---------------------------------------------
    phases = {}
    pe = stats.get("phase_events", [])
    for (n0, a), (n1, b) in zip(pe, pe[1:]):
        if n1 != "start":
            phases[n1] = phases.get(n1, 0.0) + a.elapsed_time(b) / max(1, args.steps)
    ev = stats.get("collect_events", [])
    collect_ms = sum(ev[i].elapsed_time(ev[i + 1]) for i in range(0, len(ev), 2)) / max(1, args.steps)
    n_collect = len(ev) // 2 / max(1, args.steps)
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
            _cabi.check(lib.cmh_tc_probe(engine._ptr(q_packed.sign), Q, engine._ptr(db.sign), hi - lo, BITS,
                                         engine._ptr(thr_never), tb.seg_total, tb.seg_cap, engine._ptr(tb.cand),
                                         engine._ptr(tb.cnt), engine._ptr(tb.aux), mode, engine._stream(dev)), "cmh_tc_probe")
            b.record(stream)
            torch.cuda.synchronize(dev)
            ts.append(a.elapsed_time(b))
        ceilings[name + "_ms"] = min(ts)
    del tb

    # ---- end to end through the public API with host buffers ------------------------------------------------
    from cmh_b200.synth import splitmix_rows
    q_host = _host_float_codes(SEED + 1, 0, Q).pin_memory()
    db_host = torch.empty((hi - lo, 1), dtype=torch.int64).pin_memory()
    db_host.copy_(db.sign)
    db_dev = torch.empty_like(db.sign)
    keys_host = torch.empty((Q, K), dtype=torch.int64).pin_memory()

    def e2e_step():
        # the queries go first (H2D copies share one engine: behind the shard they would wait for all of it); the
        # packed shard is uploaded in row ranges on a copy stream and the search scans each range as it lands
        qp = cu.pack_codes(q_host.to(dev, non_blocking=True), dev)
        idx = HammingIndex.from_packed_host(db_host, BITS, lo, nd_total=D, out=db_dev, pieces=3, stripes=stripes)
        k = idx.search_packed(qp, K)
        keys_host.copy_(k, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for _ in range(max(1, args.warmup - 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
    assert torch.equal(keys_host, keys.cpu()), "e2e keys differ from the device-resident run"
    h2d = q_host.numel() * 4 + db_host.numel() * 8
    d2h = keys_host.numel() * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- secondary: the reference's own evaluation calls at the shapes of configs 1-3, single GPU -----------------
    # calc_map_k_matrix in both directions (I->T, T->I), p_topK where the config names precision@N, pr_curve where it
    # names the PR curve: device float codes + labels in, pack + counting passes + host scalar out, cache cleared
    # before every call (nothing is reused between calls)
    also = None
    if not args.no_also:
        from cmh_b200.synth import CONFIGS, make_case

        def per_call_ms(fn, reps=5):
            for _ in range(2):
                cu.clear_cache(); out = fn()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for _ in range(reps):
                cu.clear_cache(); out = fn()
            torch.cuda.synchronize(dev)
            return (time.perf_counter() - t0) / reps * 1e3, out

        also = {}
        for name in ("c1", "c2-16", "c2-32", "c2-64", "c3", "c5"):     # c5: the eval stage of config 5
            shape = CONFIGS[name]
            t = make_case(shape, clustered=True, zero_query_frac=0.01)
            qi, qt, ri, rt = (torch.from_numpy(t[k]).to(dev) for k in ("q_img", "q_txt", "r_img", "r_txt"))
            qL, rL = torch.from_numpy(t["q_lab"]).to(dev), torch.from_numpy(t["r_lab"]).to(dev)
            ms_i2t, m_i2t = per_call_ms(lambda: cu.calc_map_k_matrix(qi, rt, qL, rL, shape.k, local))
            ms_t2i, m_t2i = per_call_ms(lambda: cu.calc_map_k_matrix(qt, ri, qL, rL, shape.k, local))
            entry = {"shape": f"{shape.n_query} x {shape.n_db}, {shape.bits}-bit, {shape.n_labels} labels, "
                              f"mAP@{'ALL' if shape.k is None else shape.k}",
                     "map_i2t": float(m_i2t), "map_t2i": float(m_t2i), "ms_per_call_i2t": ms_i2t, "ms_per_call_t2i": ms_t2i,
                     "map_queries_per_s": shape.n_query / (ms_i2t * 1e-3),
                     "compares_per_s": shape.n_query * shape.n_db / (ms_i2t * 1e-3)}
            if shape.topn:
                ms_p, _ = per_call_ms(lambda: cu.p_topK(qi, rt, qL, rL, list(shape.topn), local))
                entry["p_topK_ms_per_call"] = ms_p
            if name == "c3":
                ms_pr, _ = per_call_ms(lambda: cu.pr_curve(qi, rt, qL, rL, local))
                entry["pr_curve_ms_per_call"] = ms_pr
            also[name] = entry

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        nq, nd = 8, 2_000_000
        v, dt = run_reference_sample(nq, nd, 2, 1)
        cpu = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{nq} queries x {nd} rows of the same synthetic codes, 2 timed repetitions: float distance row + "
                         "full stable sort per query (utils/calc_utils.py:30-31), torch CPU, all host threads"}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
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
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "s8 (+-1 int8 tcgen05 MMA, int32 accumulate; exact integer distances and ranks)", "data": "synthetic",
        "config": config_dict(args, world),
        "e2e": {"value": Q * D / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "note": "per step: pinned-host packed database shard (uploaded in row ranges on a copy stream, scanned as "
                        "they land) + float32 query codes H2D, index build (sample), pack, tensor-core search, "
                        "(all-to-all + merge + all-gather), top-K keys D2H"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {
            "bound": "tensor", "kernel": "tc_collect_kernel (int8 tcgen05 GEMM + fused top-K candidate filter; "
                                         f"{n_collect:g} launches per step: pilot rows + the rest)",
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
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        main_reference(args)
    else:
        main_native(args)


if __name__ == "__main__":
    main()
