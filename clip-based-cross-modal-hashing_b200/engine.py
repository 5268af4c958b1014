"""Device-side objects of the retrieval-evaluation path: packed code sets and the thin wrappers that enqueue the
kernels of ``libcmh_b200.so`` on torch's current CUDA stream.

PyTorch is plumbing here (device memory, streams); every computation is a hand-written sm_100a kernel reached
through the C ABI of ``include/cmh_b200.h``.  Nothing in this module has a CPU path: tensors must live on a CUDA
device, and a missing library raises (`_cabi.lib`).

Packed words are stored in ``torch.int64`` tensors (torch has no arithmetic on uint64; the bit patterns are what
matter).  Reference sites: `utils/calc_utils.py:8-39`, `train/base.py:130-148`.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _cabi
from ._cabi import CodeSet, Plan, check

_TORCH_DTYPE = {
    torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3,
    torch.int8: 4, torch.int32: 5, torch.int64: 6, torch.uint8: 7, torch.bool: 7,
}


def _stream(device: torch.device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[ctypes.c_void_p]:
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what} must be a CUDA tensor (cmh_b200 has no CPU path)")


@dataclass
class PackedSet:
    """One side of a comparison in packed form (``cmh_codeset``).

    sign / valid : int64 [n, words]   bit c%64 of word c//64 = column c; valid is None when every entry is +-1
    labels       : int64 [n, lwords]  or None
    n_zero       : number of exact-zero code entries seen while packing (``torch.sign(0) == 0``)
    """
    sign: torch.Tensor
    valid: Optional[torch.Tensor]
    labels: Optional[torch.Tensor]
    n: int
    bits: int
    nlab: int = 0
    n_zero: int = 0

    @property
    def device(self) -> torch.device:
        return self.sign.device

    @property
    def words(self) -> int:
        return (self.bits + 63) // 64

    @property
    def lwords(self) -> int:
        return (self.nlab + 63) // 64

    def with_labels(self, labels: Optional[torch.Tensor], nlab: int) -> "PackedSet":
        return PackedSet(self.sign, self.valid, labels, self.n, self.bits, nlab if labels is not None else 0,
                         self.n_zero)

    def rows(self, lo: int, hi: int) -> "PackedSet":
        """Contiguous row range (a view) - how a database is split into shards."""
        return PackedSet(self.sign[lo:hi], None if self.valid is None else self.valid[lo:hi],
                         None if self.labels is None else self.labels[lo:hi], hi - lo, self.bits, self.nlab,
                         self.n_zero)

    def struct(self, use_valid: bool = True, use_labels: bool = True) -> CodeSet:
        cs = CodeSet()
        cs.sign = self.sign.data_ptr() if self.n else None
        cs.valid = self.valid.data_ptr() if (use_valid and self.valid is not None and self.n) else None
        cs.labels = self.labels.data_ptr() if (use_labels and self.labels is not None and self.n) else None
        cs.n = self.n
        return cs


def _as_2d(x: torch.Tensor, what: str) -> torch.Tensor:
    if x.dim() == 1:
        x = x.unsqueeze(0)
    if x.dim() != 2:
        raise ValueError(f"{what} must be 1-D or 2-D, got shape {tuple(x.shape)}")
    return x


def _row_major(x: torch.Tensor) -> torch.Tensor:
    if x.dtype not in _TORCH_DTYPE:
        raise ValueError(f"unsupported dtype {x.dtype}")
    if x.dtype == torch.bool:
        x = x.view(torch.uint8)
    if x.stride(-1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def pack_codes_device(x: torch.Tensor, counters: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Enqueue K1 on float / integer codes ``[n, bits]`` (CUDA).  Returns (sign, valid) int64 [n, words];
    ``counters`` (int64 [2], CUDA) is incremented by (#zeros, #entries outside {-1, 0, +1})."""
    _require_cuda(x, "codes")
    x = _row_major(_as_2d(x, "codes"))
    n, bits = x.shape
    if bits < 1 or bits > _cabi.CMH_MAX_BITS:
        raise ValueError(f"code length {bits} outside [1, {_cabi.CMH_MAX_BITS}]")
    words = (bits + 63) // 64
    sign = torch.empty((n, words), dtype=torch.int64, device=x.device)
    valid = torch.empty((n, words), dtype=torch.int64, device=x.device)
    with torch.cuda.device(x.device):
        check(_cabi.lib().cmh_pack_codes(_ptr(x), _TORCH_DTYPE[x.dtype], n, bits, x.stride(0) if n > 1 else bits,
                                         _ptr(sign), _ptr(valid), _ptr(counters), _stream(x.device)),
              "cmh_pack_codes")
    return sign, valid


def pack_labels_device(L: torch.Tensor, neg_counter: torch.Tensor) -> torch.Tensor:
    """Enqueue the label packer on multi-hot labels ``[n, nlab]`` (CUDA) -> int64 [n, lwords]."""
    _require_cuda(L, "labels")
    L = _row_major(_as_2d(L, "labels"))
    n, nlab = L.shape
    if nlab < 1:
        raise ValueError("labels need at least one column")
    lwords = (nlab + 63) // 64
    out = torch.empty((n, lwords), dtype=torch.int64, device=L.device)
    with torch.cuda.device(L.device):
        check(_cabi.lib().cmh_pack_labels(_ptr(L), _TORCH_DTYPE[L.dtype], n, nlab, L.stride(0) if n > 1 else nlab,
                                          _ptr(out), _ptr(neg_counter), _stream(L.device)), "cmh_pack_labels")
    return out


def synth_codes(seed: int, row0: int, n: int, bits: int, device: torch.device) -> PackedSet:
    """Counter-based packed codes generated on the device (`cmh_synth_codes`; CPU twin: `synth.splitmix_rows`)."""
    words = (bits + 63) // 64
    out = torch.empty((n, words), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        check(_cabi.lib().cmh_synth_codes(ctypes.c_uint64(seed), row0, n, bits, _ptr(out), _stream(device)),
              "cmh_synth_codes")
    return PackedSet(out, None, None, n, bits)


# ---------------------------------------------------------------------------------------------------------------
# dense blocks (a2 / a4)
# ---------------------------------------------------------------------------------------------------------------
_DENSE_MAX_ROWS = 65535 * 8


def hamming_dense(q: PackedSet, d: PackedSet) -> torch.Tensor:
    """float32 [q.n, d.n] = 0.5 * (bits - <q_i, d_j>)   (`calc_hammingDist`, utils/calc_utils.py:8-13)."""
    if q.bits != d.bits:
        raise ValueError(f"code lengths differ: {q.bits} vs {d.bits}")
    out = torch.empty((q.n, d.n), dtype=torch.float32, device=q.device)
    if q.n == 0 or d.n == 0:
        return out
    L = _cabi.lib()
    with torch.cuda.device(q.device):
        for lo in range(0, q.n, _DENSE_MAX_ROWS):
            hi = min(q.n, lo + _DENSE_MAX_ROWS)
            qs, ds = q.rows(lo, hi).struct(), d.struct()
            check(L.cmh_hamming_dense(ctypes.byref(qs), ctypes.byref(ds), q.bits, _ptr(out[lo:hi]), out.stride(0),
                                      _stream(q.device)), "cmh_hamming_dense")
    return out


def neighbor_dense(a: torch.Tensor, b: torch.Tensor, nlab: int) -> torch.Tensor:
    """float32 [na, nb] of {0, 1} from packed label masks (`calc_neighbor`, utils/calc_utils.py:42-45)."""
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    if a.shape[0] == 0 or b.shape[0] == 0:
        return out
    L = _cabi.lib()
    with torch.cuda.device(a.device):
        for lo in range(0, a.shape[0], _DENSE_MAX_ROWS):
            hi = min(a.shape[0], lo + _DENSE_MAX_ROWS)
            check(L.cmh_neighbor_dense(_ptr(a[lo:hi]), hi - lo, _ptr(b), b.shape[0], (nlab + 63) // 64,
                                       _ptr(out[lo:hi]), out.stride(0), _stream(a.device)), "cmh_neighbor_dense")
    return out


# ---------------------------------------------------------------------------------------------------------------
# ranking by counting (a3 / a5 / a6, p_topK, pr_curve, top-K)
# ---------------------------------------------------------------------------------------------------------------
class RankPass:
    """One (queries, database shard) pair walked through the two counting passes.

    Single GPU:  ``hist()`` -> ``rank(k, topn)`` -> finalisers.
    Sharded   :  ``hist()`` on every shard, exchange the shard histograms (`sharded.py`), then
                 ``rank(k, topn, lower=..., glob=...)``.
    """

    def __init__(self, q: PackedSet, d: PackedSet, *, need_labels: bool = True, max_topn: int = 0,
                 design: int = -1, ternary: Optional[bool] = None):
        if q.bits != d.bits:
            raise ValueError(f"code lengths differ: {q.bits} vs {d.bits}")
        if need_labels:
            if q.labels is None or d.labels is None:
                raise ValueError("relevance needs labels on both sides")
            if q.nlab != d.nlab:
                raise ValueError(f"label widths differ: {q.nlab} vs {d.nlab}")
        if q.device != d.device:
            raise ValueError("queries and database must be on the same device")
        self.q, self.d = q, d
        self.device = q.device
        self.need_labels = need_labels
        # sharded callers pass `ternary` so that every rank uses the same bucket layout even when only some
        # shards contain exact zeros
        self.ternary = (q.valid is not None or d.valid is not None) if ternary is None else bool(ternary)
        if not self.ternary and (q.valid is not None or d.valid is not None):
            raise ValueError("ternary=False but a valid plane is present")
        if self.ternary:
            # the ranking kernels need both planes; an all-ones plane stands in for a +-1 side
            if q.valid is None:
                q = PackedSet(q.sign, _full_valid(q), q.labels, q.n, q.bits, q.nlab)
            if d.valid is None:
                d = PackedSet(d.sign, _full_valid(d), d.labels, d.n, d.bits, d.nlab)
            self.q, self.d = q, d
        self.plan = Plan()
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_eval_plan_design(q.n, d.n, q.bits, q.nlab if need_labels else 0,
                                                   1 if self.ternary else 0, max_topn, design,
                                                   ctypes.byref(self.plan)), "cmh_eval_plan")
        self.nb = self.plan.nb
        self.workspace = torch.empty(max(1, self.plan.workspace_bytes), dtype=torch.uint8, device=self.device)
        self._qs = self.q.struct(use_labels=need_labels)
        self._ds = self.d.struct(use_labels=need_labels)
        self.hist_all: Optional[torch.Tensor] = None
        self.hist_rel: Optional[torch.Tensor] = None

    # pass 1
    def hist(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """uint32-valued int32 tensors [nq, nb]: rows of this shard per (query, bucket), all / relevant."""
        nq = self.q.n
        self.hist_all = torch.empty((nq, self.nb), dtype=torch.int32, device=self.device)
        self.hist_rel = torch.empty((nq, self.nb), dtype=torch.int32, device=self.device) if self.need_labels else None
        if nq == 0:
            return self.hist_all, self.hist_rel
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_eval_hist(ctypes.byref(self.plan), ctypes.byref(self._qs), ctypes.byref(self._ds),
                                            _ptr(self.hist_all), _ptr(self.hist_rel), _ptr(self.workspace),
                                            _stream(self.device)), "cmh_eval_hist")
        return self.hist_all, self.hist_rel

    # pass 2
    def rank(self, k: Optional[int], topn: Sequence[int] = (), lower=None, glob=None):
        """Returns (ap_sum float64 [nq], n_rel int64 [nq], hits int32 [nq, len(topn)] or None)."""
        if not self.need_labels:
            raise ValueError("rank() needs labels")
        nq = self.q.n
        ap_sum = torch.zeros(nq, dtype=torch.float64, device=self.device)
        n_rel = torch.zeros(nq, dtype=torch.int64, device=self.device)
        topn = [int(t) for t in topn]
        hits = torch.zeros((nq, len(topn)), dtype=torch.int32, device=self.device) if topn else None
        if nq == 0:
            return ap_sum, n_rel, hits
        la = lr = ga = gr = None
        if lower is not None:
            la, lr = lower
        if glob is not None:
            ga, gr = glob
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_eval_rank(ctypes.byref(self.plan), ctypes.byref(self._qs), ctypes.byref(self._ds),
                                            -1 if k is None else int(k), _ptr(la), _ptr(lr), _ptr(ga), _ptr(gr),
                                            _cabi.i64_array(topn), len(topn), _ptr(hits), _ptr(ap_sum), _ptr(n_rel),
                                            _ptr(self.workspace), _stream(self.device)), "cmh_eval_rank")
        return ap_sum, n_rel, hits

    def topk(self, K: int, index_base: int = 0) -> torch.Tensor:
        """int64 [nq, K] ascending keys ``(2*dist << 32) | (index_base + row)``; -1 (= UINT64_MAX) pads."""
        keys = torch.empty((self.q.n, int(K)), dtype=torch.int64, device=self.device)
        if self.q.n == 0:
            return keys
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_topk(ctypes.byref(self.plan), ctypes.byref(self._qs), ctypes.byref(self._ds), int(K),
                                       int(index_base), _ptr(keys), _ptr(self.workspace), _stream(self.device)),
                  "cmh_topk")
        return keys


def _full_valid(p: PackedSet) -> torch.Tensor:
    """All-ones valid plane over the real bit positions (padding bits stay 0)."""
    v = torch.full((p.n, p.words), -1, dtype=torch.int64, device=p.device)
    tail = p.bits - 64 * (p.words - 1)
    if tail < 64:
        v[:, p.words - 1] = (1 << tail) - 1
    return v


def finalize_map(ap_sum: torch.Tensor, n_rel: torch.Tensor, k: Optional[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(ap float64 [nq], mAP float32 [1]) on the device   (utils/calc_utils.py:37-38)."""
    nq = ap_sum.shape[0]
    ap = torch.empty(nq, dtype=torch.float64, device=ap_sum.device)
    out = torch.zeros(1, dtype=torch.float32, device=ap_sum.device)
    with torch.cuda.device(ap_sum.device):
        check(_cabi.lib().cmh_finalize_map(_ptr(ap_sum), _ptr(n_rel), nq, -1 if k is None else int(k), _ptr(ap),
                                           _ptr(out), _stream(ap_sum.device)), "cmh_finalize_map")
    return ap, out


def finalize_topn(hits: torch.Tensor, n_rel: torch.Tensor, topn: Sequence[int], nd_total: int) -> torch.Tensor:
    out = torch.zeros(len(topn), dtype=torch.float32, device=hits.device)
    with torch.cuda.device(hits.device):
        check(_cabi.lib().cmh_finalize_topn(_ptr(hits), _ptr(n_rel), hits.shape[0], _cabi.i64_array(topn), len(topn),
                                            int(nd_total), _ptr(out), _stream(hits.device)), "cmh_finalize_topn")
    return out


def finalize_pr(hist_all: torch.Tensor, hist_rel: torch.Tensor, bits: int, ternary: bool):
    nq = hist_all.shape[0]
    dev = hist_all.device
    P = torch.zeros(bits + 1, dtype=torch.float32, device=dev)
    R = torch.zeros(bits + 1, dtype=torch.float32, device=dev)
    L = _cabi.lib()
    ws = torch.empty(max(1, L.cmh_finalize_pr_workspace_bytes(nq, bits)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(L.cmh_finalize_pr(_ptr(hist_all), _ptr(hist_rel), nq, bits, 1 if ternary else 0, _ptr(P), _ptr(R),
                                _ptr(ws), _stream(dev)), "cmh_finalize_pr")
    return P, R


def topk_merge(keys_in: torch.Tensor, K: int) -> torch.Tensor:
    """keys_in int64 [n_lists, nq, K] (rows ascending) -> int64 [nq, K] smallest."""
    n_lists, nq, kk = keys_in.shape
    if kk != K:
        raise ValueError("keys_in last dimension must equal K")
    keys_in = keys_in.contiguous()
    out = torch.empty((nq, K), dtype=torch.int64, device=keys_in.device)
    with torch.cuda.device(keys_in.device):
        check(_cabi.lib().cmh_topk_merge(_ptr(keys_in), n_lists, nq, K, _ptr(out), _stream(keys_in.device)),
              "cmh_topk_merge")
    return out


def check_stripes(stripes, n_rows: int, index_base: int = 0) -> list:
    """Normalise the stripes of a shard: ``[(local_row, global_index), ...]`` - the local rows from ``local_row`` up to
    the next stripe (or the end of the shard) are the database rows ``global_index, global_index + 1, ...``.  A shard
    that is one contiguous row range is the single stripe ``[(0, index_base)]``."""
    if not stripes:
        return [(0, int(index_base))]
    out = [(int(lo), int(g)) for lo, g in stripes]
    if out[0][0] != 0 or any(b[0] < a[0] for a, b in zip(out, out[1:])) or out[-1][0] > max(int(n_rows), 0):
        raise ValueError("stripes must start at local row 0 and ascend within the shard")
    return out


def stripe_ranges(stripes: list, n_rows: int) -> list:
    """[(local_lo, local_hi, global_index of local_lo), ...] of the non-empty stripes."""
    ends = [lo for lo, _ in stripes[1:]] + [int(n_rows)]
    return [(lo, hi, g) for (lo, g), hi in zip(stripes, ends) if hi > lo]


def topk_exact(q: PackedSet, d: PackedSet, K: int, index_base: int = 0, stripes=None, ternary=None) -> torch.Tensor:
    """The exact two-pass (popc) top-K keys of one shard, stripe by stripe + merge when the shard has several."""
    parts = stripe_ranges(check_stripes(stripes, d.n, index_base), d.n)
    if len(parts) <= 1:
        base = parts[0][2] if parts else int(index_base)
        return RankPass(q, d.with_labels(None, 0), need_labels=False, ternary=ternary).topk(K, base)
    lists = [RankPass(q, d.rows(lo, hi).with_labels(None, 0), need_labels=False, ternary=ternary).topk(K, g)
             for lo, hi, g in parts]
    return topk_merge(torch.stack(lists), K)


# ---------------------------------------------------------------------------------------------------------------
# top-K on the tensor cores (tcgen05 / TMEM): sample histogram -> thresholds -> fused GEMM + candidate filter ->
# per-query sort; queries whose candidate list came out short or overflowed are redone by the exact two-pass path
# ---------------------------------------------------------------------------------------------------------------
TC_DEFAULT_CAP = 16384      # candidate slots per query and launch, split evenly over the launch's segments
TC_MIN_SEG = 64
TC_MAX_K = 4096
TC_PILOT_MIN_ROWS = 8_000_000     # databases at least this long get a pilot launch over their first rows
TC_PILOT_FRACTIONS = (64,)        # ... the first 1/64 of the rows, followed by a refinement of the thresholds (measured
                                  # optimum; a further stage at 1/8 gains nothing: the main launches tighten by themselves)
TC_PILOT_EARLY = 512              # shards of at least TC_PILOT_EARLY_MIN_ROWS refine once more, after 1/512 of their rows:
TC_PILOT_EARLY_MIN_ROWS = 64_000_000   # the sample's thresholds are loose (K f < 1 sample rows at the K-th distance) and
                                  # the 1/64 pilot at those costs 1.8 ms per 8192 queries on 100M rows; two stages 1.2 ms
TC_PILOT_SIGMA = 5.0
TC_PREFIX_MIN_ROWS = 4_000_000    # shards at least this long apply the prefix rule ...
TC_PREFIX_FRACTIONS = (0.3, 0.5, 0.7, 0.85)   # ... after these fractions of their rows (swept: 43.1 ms against 50.5 without)
TC_PREFIX_FRACTIONS_SHARDED = (0.3, 0.6)      # ... of every shard's rows when the database is sharded (each cut is an all-gather)


def tc_supported(q: PackedSet, d: PackedSet, K: int = 1) -> bool:
    return (q.valid is None and d.valid is None and q.bits == d.bits and 1 <= int(K) <= TC_MAX_K
            and bool(_cabi.lib().cmh_tc_supported(q.bits, 0)))


class TcBuffers:
    """Device scratch of the `cmh_tc_collect` launches over the row ranges ``regions`` of one database: candidate
    segments uint64 [nq][seg_total][seg_cap], per-segment counts uint32 [seg_total][nq], per-query bookkeeping
    uint32 [nq][8].  ``seg_base[i]`` / ``n_segs[i]``: the segments launch i fills."""

    def __init__(self, nq: int, regions: Sequence[int], bits: int, cap: int, device: torch.device,
                 seg_cap: Optional[int] = None):
        self.n_segs, self.seg_base = [], []
        with torch.cuda.device(device):
            for nd in regions:
                n = ctypes.c_int(0)
                check(_cabi.lib().cmh_tc_plan(nq, int(nd), bits, ctypes.byref(n)), "cmh_tc_plan")
                self.seg_base.append(sum(self.n_segs))
                self.n_segs.append(int(n.value))
        self.seg_total = sum(self.n_segs)
        self.n_chunks = self.seg_total
        self.seg_cap = max(TC_MIN_SEG, int(cap) // max(self.n_segs)) if seg_cap is None else int(seg_cap)
        self.cand = torch.empty((nq, self.seg_total, self.seg_cap), dtype=torch.int64, device=device)
        self.cnt = torch.empty((self.seg_total, nq), dtype=torch.int32, device=device)
        self.aux = torch.empty((nq, 8), dtype=torch.int32, device=device)
        self.thr = torch.empty(nq, dtype=torch.int32, device=device)
        self.thr2 = torch.empty(nq, dtype=torch.int32, device=device)
        self.thr3 = torch.empty(nq, dtype=torch.int32, device=device)
        self.thr4 = torch.empty(nq, dtype=torch.int32, device=device)
        self.fail_flags = torch.empty(nq, dtype=torch.int32, device=device)
        self.fail_count = torch.zeros(1, dtype=torch.int32, device=device)


def tc_pilot_rows(nd: int) -> int:
    """Rows scanned by the pilot launches (0 = none): a multiple of the 256-row tile."""
    if nd < TC_PILOT_MIN_ROWS:
        return 0
    return (nd // TC_PILOT_FRACTIONS[-1]) // 256 * 256


def tc_pilot_stages(nd: int, nd_total: int, world: int = 1) -> list:
    """Cumulative row counts (multiples of the 256-row tile, ascending, < nd) after which the thresholds are refined.
    The NUMBER of stages depends on the whole database and the number of shards only - every shard takes part in
    every refinement."""
    if nd_total < TC_PILOT_MIN_ROWS:
        return []
    fractions = TC_PILOT_FRACTIONS
    if nd_total // max(1, int(world)) >= TC_PILOT_EARLY_MIN_ROWS:
        fractions = (TC_PILOT_EARLY,) + tuple(fractions)
    return [(nd // f) // 256 * 256 for f in fractions]


class LocalComm:
    """The exchange steps of the tensor-core top-K for a database that lives on ONE GPU (no-ops).  `sharded.GroupComm`
    is the `torch.distributed` version for a database sharded over the ranks of a process group."""
    world = 1
    rank = 0

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        return t

    def all_gather_stack(self, t: torch.Tensor) -> torch.Tensor:
        return t.unsqueeze(0)

    def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
        return t


def topk_tc(q: PackedSet, d: PackedSet, K: int, index_base: int = 0, sample: Optional[PackedSet] = None,
            cap: int = TC_DEFAULT_CAP, stats: Optional[dict] = None, tighten: bool = True,
            seg_cap: Optional[int] = None, pilot: Optional[int] = None, comm=None, nd_total: Optional[int] = None,
            exact_fallback=None, buffers: Optional[dict] = None, ready=None, defer: bool = False, prefix: bool = True,
            stripes=None):
    """int64 [nq, K] ascending keys, identical to ``RankPass(q, d).topk(K, index_base)`` (to the global stable
    ranking when ``d`` is one shard of a database of ``nd_total`` rows and ``comm`` spans the shards).

    sample  a subset of the rows of ``d`` (any rows, contiguous in memory) used only to guess the per-query
            thresholds; None = use ``d`` itself (exact thresholds when unsharded, an extra popc pass)
    pilot   cumulative row counts of the pilot launches (None = `tc_pilot_stages`; an int = one stage): the rows up
            to each count are scanned with the thresholds known so far, and what they hold refines the thresholds
            for the rest
    comm    exchange steps (`LocalComm`, `sharded.GroupComm`): the sample / pilot histograms are all-reduced so that
            every shard filters with the same global thresholds - a shard then contributes only its share of the ~K
            rows below them; the per-shard results are exchanged all-to-all (rank r merges the r-th slice of the
            queries: 1/N of the traffic and of the merge work of an all-gather) and the merged slices all-gathered
    ready   [(row_end, torch.cuda.Event), ...] in row order: rows below row_end of ``d`` are valid once the event has
            completed (a database that is still being uploaded on another stream); the scan is cut at those
            boundaries and every launch waits only for the rows it reads
    prefix  apply the prefix rule (`cmh_tc_choose_prefix` / `cmh_tc_choose_seen`) at `TC_PREFIX_FRACTIONS` of the rows:
            exact; sharded databases all-gather the candidate histograms of the rows scanned so far at each cut
    stripes ``[(local_row, global_index), ...]`` (`check_stripes`): the shard is several row ranges of the database
            instead of one (``index_base`` is then ignored).  With ``comm`` spanning several shards the stripes must be
            LOCKSTEP stripes - the same number on every shard, stripe j of every shard below stripe j+1 of every shard
            in global index: the prefix rule is then applied at the stripe boundaries on the all-reduced histograms,
            with everything scanned so far (on any shard) of lower index than everything still to come, as on one GPU
    defer   return a callable instead of the keys: everything is enqueued, and calling it reads the verdict (a host
            sync), redoes failed queries and returns the keys - lets a caller keep two query chunks in flight
    buffers a dict the caller keeps between calls: the multi-GB candidate scratch is allocated once per query-chunk
            size instead of per call
    exact_fallback(sub_q) -> keys for the queries whose candidate lists came out short or overflowed (the exact
            two-pass path; default: `RankPass.topk` on ``d``, which is only right when unsharded)"""
    if not tc_supported(q, d, K):
        raise ValueError("tensor-core top-K needs +-1 codes of 64 or 128 bits and K <= 4096")
    comm = LocalComm() if comm is None else comm
    K = int(K)
    dev = q.device
    nq = q.n
    nd_total = d.n if nd_total is None else int(nd_total)
    per_rank = -(-nq // comm.world)                  # queries merged by each rank (the last slice may be padded)
    keys_all = torch.empty((per_rank * comm.world, K), dtype=torch.int64, device=dev)
    keys = keys_all[:nq]
    if nq == 0 or nd_total == 0:
        if nq and nd_total == 0:
            keys.fill_(-1)
        done = keys
        return (lambda: done) if defer else done
    if keys_all.shape[0] > nq:
        keys_all[nq:].fill_(-1)
    L = _cabi.lib()
    nb = q.bits + 1
    smp = d if sample is None else sample
    # refinement stages: cumulative local row counts; the same number of stages on every shard
    if pilot is None:
        stages = tc_pilot_stages(d.n, nd_total, comm.world)
    else:
        stages = [int(x) for x in (pilot if isinstance(pilot, (list, tuple)) else [pilot]) if int(x) > 0]
    if sample is None and comm.world == 1:
        stages = []                                  # exact thresholds need no refinement
    stages = [min(e, d.n) for e in stages]
    ends = sorted({e for e in stages if 0 < e < d.n})
    # the prefix rule (exact): after these rows the candidates so far bound what later rows can still contribute
    # (every shard takes part in every exchange, so whether and how often is decided from the global sizes alone)
    stripes = check_stripes(stripes, d.n, index_base)
    lockstep = comm.world > 1 and len(stripes) > 1
    ascending = all(b[1] >= a[1] + (b[0] - a[0]) for a, b in zip(stripes, stripes[1:]))
    prefix_cuts = []
    if lockstep:
        if prefix:
            prefix_cuts = [lo for lo, _ in stripes[1:]]
    elif prefix and ascending and -(-nd_total // comm.world) >= TC_PREFIX_MIN_ROWS:
        fractions = TC_PREFIX_FRACTIONS if comm.world == 1 else TC_PREFIX_FRACTIONS_SHARDED
        prefix_cuts = [min(d.n, int(d.n * f) // 256 * 256) for f in fractions]
    prefix_ends = {e for e in prefix_cuts if (stages[-1] if stages else 0) < e < d.n}
    cuts = sorted({0, d.n} | set(ends) | prefix_ends | {int(e) for e, _ in (ready or ()) if 0 < int(e) < d.n} |
                  {lo for lo, _ in stripes if 0 < lo < d.n})

    def global_index(row: int) -> int:               # of a local row (spans never straddle a stripe boundary)
        lo, g = [st_ for st_ in stripes if st_[0] <= row][-1]
        return g + (row - lo)
    spans = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]] or [(0, 0)]
    regions = [hi - lo for lo, hi in spans]

    def wait_rows(hi: int) -> None:
        for end, ev in (ready or ()):
            if int(end) >= hi:
                torch.cuda.current_stream(dev).wait_event(ev)
                return

    bkey = (nq, tuple(regions), q.bits, int(cap), seg_cap, str(dev))
    b = buffers.get(bkey) if buffers is not None else None
    if b is None:
        b = TcBuffers(nq, regions, q.bits, cap, dev, seg_cap)
        if buffers is not None:
            buffers.clear()                          # one geometry at a time: the scratch is large
            buffers[bkey] = b
    if comm.world > 1:
        ckey = ("counts", smp.n, tuple(stages))
        got = buffers.get(ckey) if buffers is not None else None
        if got is None:
            counts = comm.all_reduce_sum(torch.tensor([smp.n] + stages, dtype=torch.int64, device=dev))
            got = tuple(int(v) for v in counts.tolist())
            if buffers is not None:
                buffers[ckey] = got
        n_sample_all, stages_all = got[0], list(got[1:])
    else:
        n_sample_all, stages_all = smp.n, list(stages)
    if stats is not None and stats.get("time_phases"):
        e = torch.cuda.Event(enable_timing=True)
        e.record(torch.cuda.current_stream(dev))
        stats.setdefault("phase_events", []).append(("start", e))
    if smp.n:
        h_all, _ = RankPass(q.with_labels(None, 0), smp.with_labels(None, 0), need_labels=False).hist()
    else:
        h_all = torch.zeros((nq, nb), dtype=torch.int32, device=dev)
    h_all = comm.all_reduce_sum(h_all)
    with torch.cuda.device(dev):
        st = _stream(dev)
        check(L.cmh_topk_threshold(_ptr(h_all), nq, nb, n_sample_all, nd_total, K, _ptr(b.thr), st), "cmh_topk_threshold")
        timed = stats is not None and stats.get("time_collect")

        def mark():
            if timed:
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(dev))
                stats.setdefault("collect_events", []).append(e)

        def phase(name):
            if stats is not None and stats.get("time_phases"):
                e = torch.cuda.Event(enable_timing=True)
                e.record(torch.cuda.current_stream(dev))
                stats.setdefault("phase_events", []).append((name, e))

        phase("thresholds_done")
        thr_cur, thr_next = b.thr, b.thr2
        thr_limit = None
        last_stage = stages[-1] if stages else 0
        si = 0                                       # next stage to close
        launched = False

        def refine_upto(i_stage: int, seg_hi: int):
            # the candidates of the rows scanned so far (kept at thresholds >= the current ones): exact counts of
            # every bucket at or below the current threshold; all-reduced over the shards
            nonlocal thr_cur, thr_next
            ph = torch.zeros((nq, nb), dtype=torch.int32, device=dev)
            over = torch.zeros(nq, dtype=torch.int32, device=dev)
            if seg_hi > 0:
                check(L.cmh_tc_cand_hist(_ptr(b.cand), _ptr(b.cnt), nq, 0, seg_hi, b.seg_total, b.seg_cap, nb, _ptr(ph),
                                         _ptr(over), st), "cmh_tc_cand_hist")
            ph, over = comm.all_reduce_sum(ph), comm.all_reduce_max(over)
            check(L.cmh_tc_choose(_ptr(ph), _ptr(over), nq, nb, stages_all[i_stage], nd_total, K, TC_PILOT_SIGMA,
                                  _ptr(thr_cur), _ptr(thr_next), st), "cmh_tc_choose")
            thr_cur, thr_next = thr_next, thr_cur

        def prefix_upto(seg_hi: int):
            # The prefix rule (exact, no statistics).  K candidates at dist <= b among rows of LOWER index than what is
            # still to be scanned (this shard's rows so far + the same prefix of every lower-ranked shard): later rows
            # only matter below b.  K candidates at dist <= b ANYWHERE among the rows scanned so far (all shards):
            # later rows only matter at or below b.  One all-gather of the per-shard histograms serves both.
            nonlocal thr_cur, thr_limit
            if thr_limit is None:
                thr_limit = thr_cur                  # the statistical bound, the same on every shard
            ph = torch.zeros((nq, nb), dtype=torch.int32, device=dev)
            over = torch.zeros(nq, dtype=torch.int32, device=dev)
            if seg_hi > 0:
                check(L.cmh_tc_cand_hist(_ptr(b.cand), _ptr(b.cnt), nq, 0, seg_hi, b.seg_total, b.seg_cap, nb, _ptr(ph),
                                         _ptr(over), st), "cmh_tc_cand_hist")
            out = b.thr3 if thr_cur is not b.thr3 else b.thr4
            if lockstep:
                # lockstep stripes: whatever any shard has scanned lies below whatever any shard has still to scan
                ph = comm.all_reduce_sum(ph)
                check(L.cmh_tc_choose_prefix(_ptr(ph), None, nq, nb, K, _ptr(thr_cur), _ptr(out), st), "cmh_tc_choose_prefix")
            elif comm.world > 1:
                every = comm.all_gather_stack(ph)                                    # [world, nq, nb]
                lower = every[:comm.rank + 1].sum(0, dtype=torch.int32).contiguous()
                seen = every.sum(0, dtype=torch.int32).contiguous()
                check(L.cmh_tc_choose_prefix(_ptr(lower), None, nq, nb, K, _ptr(thr_cur), _ptr(out), st), "cmh_tc_choose_prefix")
                check(L.cmh_tc_choose_seen(_ptr(seen), None, nq, nb, K, _ptr(out), _ptr(out), st), "cmh_tc_choose_seen")
            else:
                check(L.cmh_tc_choose_prefix(_ptr(ph), _ptr(over), nq, nb, K, _ptr(thr_cur), _ptr(out), st), "cmh_tc_choose_prefix")
            thr_cur = out

        pj = 0                                       # next prefix exchange

        def close(hi: int, seg_hi: int):
            # exchanges due once the rows below `hi` have been scanned: the pilot stages, then - never before the last
            # stage, so that the order is the same on every shard - the prefix rule.  A shard that has no rows at a
            # cut (a short or empty shard) still takes part.
            nonlocal si, pj
            while si < len(stages) and stages[si] <= hi:
                if stages_all[si] > 0:
                    refine_upto(si, seg_hi)
                si += 1
                phase("pilot_done")
            while si == len(stages) and pj < len(prefix_cuts) and prefix_cuts[pj] <= hi:
                prefix_upto(seg_hi)
                pj += 1

        close(0, 0)
        for i, (lo, hi) in enumerate(spans):
            if hi <= lo:
                continue
            in_pilot = hi <= last_stage              # pilot spans keep EVERY row at or below the threshold (K = 0)
            if not in_pilot and thr_limit is None:
                thr_limit = thr_cur                  # the statistical bound; later (prefix) thresholds are exact exclusions
            wait_rows(hi)
            mark()
            # tightening (main spans) uses this launch's own counts: K rows found locally are K rows found globally
            check(L.cmh_tc_collect(_ptr(q.sign), nq, _ptr(d.sign[lo:]), hi - lo, q.bits, global_index(lo),
                                   _ptr(thr_cur), 0 if in_pilot or not tighten else K, b.seg_base[i], b.seg_total,
                                   b.seg_cap, _ptr(b.cand), _ptr(b.cnt), _ptr(b.aux), st), "cmh_tc_collect")
            mark()
            launched = True
            if hi < d.n:                             # after the last row there is nothing left to tighten for
                close(hi, b.seg_base[i] + b.n_segs[i])
        close(d.n, b.seg_total if launched else 0)
        if not launched:
            b.cnt.zero_()
            b.aux.zero_()
        thr_main = thr_limit if thr_limit is not None else thr_cur
        phase("main_done")
        partial = 1 if comm.world > 1 else 0
        check(L.cmh_topk_finalize(_ptr(b.cand), _ptr(b.cnt), _ptr(b.aux), _ptr(thr_main), nq, b.seg_total, b.seg_cap, K,
                                  nd_total, partial, _ptr(keys), _ptr(b.fail_flags), _ptr(b.fail_count), st),
              "cmh_topk_finalize")
        phase("finalize_done")
        if comm.world > 1:
            mine = topk_merge(comm.all_to_all(keys_all.view(comm.world, per_rank, K)), K)     # [per_rank, K]
            keys = comm.all_gather_stack(mine).view(per_rank * comm.world, K)[:nq]
            b.fail_flags = comm.all_reduce_max(b.fail_flags)
            check(L.cmh_topk_verify(_ptr(keys), _ptr(thr_main), nq, K, nd_total, _ptr(b.fail_flags), _ptr(b.fail_count),
                                    st), "cmh_topk_verify")
    phase("exchange_done")
    # the verdict is read by `finish`: at once, or - `defer` - when the caller asks for the result, so that the next
    # query chunk can be enqueued (on another stream) before this one has drained
    fail_count = b.fail_count.clone() if defer else b.fail_count
    fail_flags = b.fail_flags.clone() if defer else b.fail_flags
    if stats is not None:
        stats["candidates"] = b.cnt.sum(0)
        stats["thr"] = thr_main
        stats["thr_final"] = thr_cur                 # after the prefix rule (differs between shards)
        stats["pilot_rows"] = stages

    def finish() -> torch.Tensor:
        n_fail = int(fail_count.item())
        if stats is not None:
            stats["n_fail"] = n_fail
        if n_fail:
            rows = torch.nonzero(fail_flags, as_tuple=False).squeeze(1)
            sub = PackedSet(q.sign.index_select(0, rows).contiguous(), None, None, int(rows.numel()), q.bits)
            if exact_fallback is None:
                if comm.world > 1:
                    raise RuntimeError("a sharded tensor-core top-K needs exact_fallback")
                redo = topk_exact(sub, d, K, index_base, stripes)
            else:
                redo = exact_fallback(sub)
            keys.index_copy_(0, rows, redo)
        return keys

    return finish if defer else finish()
