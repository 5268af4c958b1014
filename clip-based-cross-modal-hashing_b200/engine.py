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


def pack_codes_device(x: torch.Tensor, counters: torch.Tensor, device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Enqueue K1 on float / integer codes ``[n, bits]``.  Returns (sign, valid) int64 [n, words] on the device;
    ``counters`` (int64 [2], CUDA) is incremented by (#zeros, #entries outside {-1, 0, +1}).

    ``x`` is a CUDA tensor - or a PINNED host tensor (``device`` then names the GPU): page-locked memory is mapped into the
    device's address space, so the kernel reads the codes straight over the link and only the packed words (1/32 of the
    bytes for float32) ever exist in HBM; no staging copy, nothing queued on the copy engines."""
    if not x.is_cuda:
        if not x.is_pinned() or device is None:
            raise RuntimeError("codes must be a CUDA tensor or a pinned host tensor with a target device "
                               "(cmh_b200 has no CPU path)")
        dev = torch.device(device)
    else:
        dev = x.device
    x = _row_major(_as_2d(x, "codes"))
    if not x.is_cuda and not x.is_pinned():      # (a layout fix-up made a pageable copy)
        x = x.to(dev)
    n, bits = x.shape
    if bits < 1 or bits > _cabi.CMH_MAX_BITS:
        raise ValueError(f"code length {bits} outside [1, {_cabi.CMH_MAX_BITS}]")
    words = (bits + 63) // 64
    sign = torch.empty((n, words), dtype=torch.int64, device=dev)
    valid = torch.empty((n, words), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        check(_cabi.lib().cmh_pack_codes(_ptr(x), _TORCH_DTYPE[x.dtype], n, bits, x.stride(0) if n > 1 else bits,
                                         _ptr(sign), _ptr(valid), _ptr(counters), _stream(dev)),
              "cmh_pack_codes")
    if not x.is_cuda:
        sign._cmh_keepalive = x                  # the kernel is still reading the host pages when this returns
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


def tc_supported(q: PackedSet, d: PackedSet, K: int = 1) -> bool:
    return (q.valid is None and d.valid is None and q.bits == d.bits and 1 <= int(K) <= TC_MAX_K
            and bool(_cabi.lib().cmh_tc_supported(q.bits, 0)))


class TcBuffers:
    """Device scratch of bare `cmh_tc_collect` launches (measurement aids and kernel-level tests; a search owns its
    scratch through `cmh_tc_search_plan`): candidate segments uint64 [nq][seg_total][seg_cap], per-segment counts
    uint32 [seg_total][nq], per-query bookkeeping uint32 [nq][8]."""

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


def tc_pilot_stages(nd: int, nd_total: int, world: int = 1) -> list:
    """Cumulative row counts (multiples of the 256-row tile) after which a search refines its thresholds
    (`cmh_tc_pilot_stages`).  The NUMBER of stages depends on the whole database and the number of shards only."""
    rows = (ctypes.c_int64 * _cabi.TC_MAX_STAGES)()
    n = _cabi.lib().cmh_tc_pilot_stages(int(nd), int(nd_total), int(world), rows)
    return [int(rows[i]) for i in range(n)]


def tc_pilot_rows(nd: int) -> int:
    """Rows scanned by the pilot launches of an unsharded database (0 = none)."""
    st = tc_pilot_stages(nd, nd, 1)
    return st[-1] if st else 0


class LocalComm:
    """A database that lives on ONE GPU: no exchange steps.  `sharded.GroupComm` spans the ranks of a
    `torch.distributed` process group (NCCL transport inside the library)."""
    world = 1
    rank = 0

    def handle(self):
        return None


class _DevBytes:
    """A device pointer as something `torch.as_tensor` accepts (callback transports only)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class CallbackComm:
    """`cmh_comm` filled with Python functions: wraps any object with ``rank``, ``world``, ``all_reduce_sum(t)``,
    ``all_reduce_max(t)``, ``all_gather_stack(t)`` and ``all_to_all(t)`` (tensors in, tensors out) as the library's
    transport.  This is the bring-your-own-transport form of the C ABI; the tests drive several shards of one GPU
    through it with a barrier-based double.  The functions run on the caller's thread, on torch's current stream."""

    def __init__(self, comm, device: torch.device):
        self.comm, self.device, self.error = comm, device, None
        self.rank, self.world = int(comm.rank), int(comm.world)

        def view(ptr, nbytes, dtype=torch.uint8):
            return torch.as_tensor(_DevBytes(ptr, nbytes), device=self.device).view(dtype)

        def guard(fn):
            def run(*a):
                try:
                    fn(*a)
                    return 0
                except BaseException as e:  # noqa: BLE001 - reported through the C ABI's return code
                    self.error = e
                    return 1
            return run

        def all_reduce(_ctx, buf, count, op, _stream):
            t = view(buf, count * 4, torch.int32)
            t.copy_(self.comm.all_reduce_max(t.clone()) if op == 1 else self.comm.all_reduce_sum(t.clone()))

        def all_gather(_ctx, send, recv, nbytes, _stream):
            got = self.comm.all_gather_stack(view(send, nbytes).clone())
            view(recv, nbytes * self.world).copy_(got.reshape(-1))

        def all_to_all(_ctx, send, recv, nbytes, _stream):
            got = self.comm.all_to_all(view(send, nbytes * self.world).clone().view(self.world, nbytes))
            view(recv, nbytes * self.world).copy_(got.reshape(-1))

        self._fns = (_cabi.COMM_ALL_REDUCE(guard(all_reduce)), _cabi.COMM_ALL_GATHER(guard(all_gather)),
                     _cabi.COMM_ALL_TO_ALL(guard(all_to_all)))
        self.struct = _cabi.Comm(None, self.rank, self.world, *self._fns)

    def handle(self):
        return ctypes.pointer(self.struct)


def _comm_handle(comm, device):
    """(ctypes pointer to a `cmh_comm` or None, the object that must stay alive while it is used)"""
    if comm is None:
        return None, None
    if hasattr(comm, "handle"):
        return comm.handle(), comm
    cb = CallbackComm(comm, device)
    return cb.handle(), cb


# Verdicts travel to pinned host memory behind their own search and are waited for by EVENT: a stream-ordered read
# (`.item()`) would queue behind whatever was enqueued after the search - e.g. the next query chunk.  One process-wide
# ring (pinning memory costs a driver call; an index rebuilt every step must not pay it again).
_VERDICTS: list = []
_verdict_turn = 0


def _verdict_slot():
    global _verdict_turn
    if not _VERDICTS:
        host = torch.empty(16, dtype=torch.int32).pin_memory()
        _VERDICTS.extend((host[i:i + 1], torch.cuda.Event()) for i in range(16))
    slot = _VERDICTS[_verdict_turn % len(_VERDICTS)]
    _verdict_turn += 1
    return slot


class TcSearchPlan:
    """`cmh_tc_search` + its device scratch (+ optional timing handle) for one (queries per call, shard) geometry."""

    def __init__(self, plan: "_cabi.TcSearch", device: torch.device):
        self.plan = plan
        self.device = device
        self.workspace = torch.empty(max(1, int(plan.workspace_bytes)), dtype=torch.uint8, device=device)
        self._timing = None

    def verdict_slot(self):
        return _verdict_slot()

    def timing(self):
        if self._timing is None:
            h = ctypes.c_void_p()
            with torch.cuda.device(self.device):
                check(_cabi.lib().cmh_tc_timing_create(ctypes.byref(h)), "cmh_tc_timing_create")
            self._timing = h
        return self._timing

    def read_timing(self):
        ph = (ctypes.c_float * _cabi.TC_PHASES)()
        col, n = ctypes.c_float(0), ctypes.c_int(0)
        check(_cabi.lib().cmh_tc_timing_read(self._timing, ph, ctypes.byref(col), ctypes.byref(n)), "cmh_tc_timing_read")
        return {name: float(ph[i]) for i, name in enumerate(_cabi.TC_PHASE_NAMES)}, float(col.value), int(n.value)

    def launch_ms(self) -> list:
        """Device time of each tc_collect launch of the last timed search."""
        ms = (ctypes.c_float * _cabi.TC_MAX_SPANS)()
        check(_cabi.lib().cmh_tc_timing_launches(self._timing, ms, _cabi.TC_MAX_SPANS), "cmh_tc_timing_launches")
        n = sum(1 for i in range(self.plan.n_spans) if self.plan.span_hi[i] > self.plan.span_lo[i])
        return [float(ms[i]) for i in range(n)]

    def _view(self, off: int, shape, dtype):
        n = 1
        for v in shape:
            n *= int(v)
        return self.workspace[off:off + n * dtype.itemsize].view(dtype).view(*shape)

    @property
    def cnt(self) -> torch.Tensor:
        return self._view(int(self.plan.off_cnt), (self.plan.seg_total, self.plan.nq), torch.int32)

    def thr(self, slot: int) -> torch.Tensor:
        return self._view(int(self.plan.off_thr) + int(slot) * int(self.plan.nq) * 4, (self.plan.nq,), torch.int32)

    def __del__(self):
        if getattr(self, "_timing", None) is not None:
            try:
                _cabi.lib().cmh_tc_timing_destroy(self._timing)
            except Exception:  # noqa: BLE001 - interpreter shutdown
                pass


def tc_search_plan(comm_handle, nq: int, nd: int, nd_total: int, bits: int, K: int, stripes, n_sample: int, *,
                   pilot=None, prefix: bool = True, prefix_fractions=None, prefix_min_rows: Optional[int] = None,
                   tighten: bool = True, cap: int = TC_DEFAULT_CAP, seg_cap: Optional[int] = None,
                   exact_thresholds: bool = False, gather: bool = True, ready_rows=(), device=None) -> "_cabi.TcSearch":
    """`cmh_tc_search_plan` (collective over the comm's ranks when it spans several)."""
    L = _cabi.lib()
    o = _cabi.TcOpts()
    L.cmh_tc_default_opts(ctypes.byref(o))
    if pilot is not None:
        rows = [int(x) for x in (pilot if isinstance(pilot, (list, tuple)) else [pilot]) if int(x) > 0]
        if len(rows) > _cabi.TC_MAX_STAGES:
            raise ValueError(f"at most {_cabi.TC_MAX_STAGES} pilot stages")
        o.n_pilot = len(rows)
        for i, r in enumerate(rows):
            o.pilot_rows[i] = r
    o.prefix = 1 if prefix else 0
    if prefix_fractions is not None:
        fr = [float(f) for f in prefix_fractions]
        if len(fr) > _cabi.TC_MAX_CUTS:
            raise ValueError(f"at most {_cabi.TC_MAX_CUTS} prefix cuts")
        o.n_prefix = len(fr)
        for i, f in enumerate(fr):
            o.prefix_frac[i] = f
    if prefix_min_rows is not None:
        o.prefix_min_rows = int(prefix_min_rows)
    o.tighten = 1 if tighten else 0
    o.cap = int(cap)
    o.seg_cap = 0 if seg_cap is None else int(seg_cap)
    o.exact_thresholds = 1 if exact_thresholds else 0
    o.gather = 1 if gather else 0
    ready_rows = [int(r) for r in ready_rows]
    if len(ready_rows) > _cabi.TC_MAX_READY:
        raise ValueError(f"at most {_cabi.TC_MAX_READY} upload ranges")
    o.n_ready = len(ready_rows)
    for i, r in enumerate(ready_rows):
        o.ready_rows[i] = r
    srow = (ctypes.c_int64 * len(stripes))(*[int(a) for a, _ in stripes])
    sidx = (ctypes.c_int64 * len(stripes))(*[int(b) for _, b in stripes])
    plan = _cabi.TcSearch()
    import contextlib
    with (torch.cuda.device(device) if device is not None else contextlib.nullcontext()):   # host arithmetic on one GPU
        check(L.cmh_tc_search_plan(comm_handle, int(nq), int(nd), int(nd_total), int(bits), int(K), len(stripes), srow, sidx,
                                   int(n_sample), ctypes.byref(o), ctypes.byref(plan)), "cmh_tc_search_plan")
    return plan


def topk_tc(q: PackedSet, d: PackedSet, K: int, index_base: int = 0, sample: Optional[PackedSet] = None,
            cap: int = TC_DEFAULT_CAP, stats: Optional[dict] = None, tighten: bool = True,
            seg_cap: Optional[int] = None, pilot=None, comm=None, nd_total: Optional[int] = None,
            exact_fallback=None, buffers: Optional[dict] = None, ready=None, defer: bool = False, prefix: bool = True,
            stripes=None, gather: bool = True, prefix_fractions=None, prefix_min_rows: Optional[int] = None):
    """Top-``K`` keys on the tensor cores - `cmh_topk_tc`, one library call that owns the whole launch chain (sample
    histogram -> thresholds -> pilot launches + refinement -> main launches with the exact prefix rule -> finalize ->
    exchange).  int64 [nq, K] ascending keys, identical to ``RankPass(q, d).topk(K, index_base)`` (to the global stable
    ranking when ``d`` is one shard of a database of ``nd_total`` rows and ``comm`` spans the shards).

    sample  a subset of the rows of ``d`` (any rows, contiguous in memory) used only to guess the per-query
            thresholds; None = a full popc histogram of ``d`` (exact thresholds; one GPU only)
    pilot   cumulative row counts of the pilot launches (None = `tc_pilot_stages`; an int = one stage)
    comm    `LocalComm` / `sharded.GroupComm` (NCCL inside the library) / any object with the exchange methods of
            `CallbackComm`: every shard filters with the same global thresholds and contributes only its share of
            the ~K rows below them; rank r merges and verifies the r-th slice of the queries
    gather  sharded: True = every rank returns all [nq, K] keys; False = rank r returns ITS slice, int64 [per_rank, K] -
            the keys of queries ``r * per_rank ...`` (per_rank = ceil(nq / world); rows past nq are pads)
    ready   [(row_end, torch.cuda.Event), ...] in row order: rows below row_end of ``d`` are valid once the event has
            completed (a database that is still being uploaded on another stream)
    prefix  apply the exact prefix rule (at ``prefix_fractions`` of the shard's rows, default 0.3 / 0.5 / 0.7 / 0.85 on
            one GPU, 0.3 / 0.6 for contiguous shards; at the stripe boundaries for lockstep stripes)
    stripes ``[(local_row, global_index), ...]`` (`check_stripes`); with a comm of several ranks they must be LOCKSTEP
            stripes (`sharded.lockstep_stripes`)
    defer   return a callable instead of the keys: everything is enqueued, and calling it reads the verdict (a host
            sync), redoes failed queries and returns the keys
    buffers a dict the caller keeps between calls: plan + multi-GB candidate scratch are made once per geometry
    exact_fallback(sub_q) -> [n, K] global keys of the queries whose candidate lists came out short or overflowed"""
    if not tc_supported(q, d, K):
        raise ValueError("tensor-core top-K needs +-1 codes of at most 128 bits and K <= 4096")
    K = int(K)
    dev = q.device
    nq = q.n
    world = 1 if comm is None else int(comm.world)
    rank = 0 if comm is None else int(comm.rank)
    nd_total = d.n if nd_total is None else int(nd_total)
    per_rank = -(-nq // world)
    sliced = world > 1 and not gather
    if nq == 0 or nd_total == 0:
        done = torch.full((per_rank if sliced else nq, K), -1, dtype=torch.int64, device=dev)
        return (lambda: done) if defer else done
    stripes = check_stripes(stripes, d.n, index_base)
    exact_thr = sample is None
    if exact_thr and world > 1:
        raise ValueError("a sharded tensor-core top-K needs a sample of every shard")
    n_sample = 0 if exact_thr else sample.n
    ready = list(ready or ())
    handle, keep = _comm_handle(comm, dev)
    key = (nq, d.n, nd_total, q.bits, K, tuple(stripes), n_sample, exact_thr,
           None if pilot is None else tuple(pilot) if isinstance(pilot, (list, tuple)) else int(pilot), bool(prefix),
           None if prefix_fractions is None else tuple(prefix_fractions), prefix_min_rows, bool(tighten), int(cap), seg_cap,
           bool(gather), tuple(int(e) for e, _ in ready), world, rank, str(dev),
           torch.cuda.current_stream(dev).cuda_stream)       # (scratch is only ever shared by searches of ONE stream)
    sp = buffers.get(key) if buffers is not None else None
    if sp is None:
        plan = tc_search_plan(handle, nq, d.n, nd_total, q.bits, K, stripes, n_sample, pilot=pilot, prefix=prefix,
                              prefix_fractions=prefix_fractions, prefix_min_rows=prefix_min_rows, tighten=tighten, cap=cap,
                              seg_cap=seg_cap, exact_thresholds=exact_thr, gather=gather,
                              ready_rows=[e for e, _ in ready], device=dev)
        if buffers is not None:
            while len(buffers) >= 2:                 # at most two geometries at a time: the scratch is large
                buffers.pop(next(iter(buffers)))
        sp = TcSearchPlan(plan, dev)
        if buffers is not None:
            buffers[key] = sp
    p = sp.plan
    keys = torch.empty((per_rank if sliced else (per_rank * world if world > 1 else nq), K), dtype=torch.int64, device=dev)
    fail_flags = torch.empty(per_rank * world, dtype=torch.int32, device=dev)
    fail_count = torch.empty(1, dtype=torch.int32, device=dev)
    timed = stats is not None and (stats.get("time_phases") or stats.get("time_collect"))
    events = None
    if ready:
        events = (ctypes.c_void_p * len(ready))(*[ctypes.c_void_p(ev.cuda_event) for _, ev in ready])
    with torch.cuda.device(dev):
        rc = _cabi.lib().cmh_topk_tc(ctypes.byref(p), handle, _ptr(q.sign), _ptr(d.sign) if d.n else None,
                                     None if exact_thr or n_sample == 0 else _ptr(sample.sign), events, _ptr(keys),
                                     _ptr(fail_flags), _ptr(fail_count), _ptr(sp.workspace),
                                     sp.timing() if timed else None, _stream(dev))
    if rc and getattr(keep, "error", None) is not None:
        raise keep.error
    check(rc, "cmh_topk_tc")
    fail_host, fail_event = sp.verdict_slot()
    fail_host.copy_(fail_count, non_blocking=True)
    fail_event.record(torch.cuda.current_stream(dev))
    if stats is not None:
        stats["candidates"] = sp.cnt.sum(0)
        stats["thr"] = sp.thr(p.thr_limit_slot).clone()
        stats["thr_final"] = sp.thr(p.thr_final_slot).clone()      # after the prefix rule (differs between contiguous shards)
        stats["pilot_rows"] = [int(p.stage_rows[i]) for i in range(p.n_stages)]
        stats["n_launches"] = sum(1 for i in range(p.n_spans) if p.span_hi[i] > p.span_lo[i])
        stats["exch_width"] = int(p.exch_width)

    verdict: dict = {}

    def finish() -> torch.Tensor:
        fail_event.synchronize()                     # the one host sync of a search: this search's verdict, nothing later
        n_fail = int(fail_host[0])
        verdict["n_fail"] = n_fail                   # (a deferred caller can tell whether redo work was enqueued)
        if stats is not None:
            stats["n_fail"] = n_fail
            if timed:
                ph, col, n_col = sp.read_timing()
                acc = stats.setdefault("phase_ms_sum", {})
                for name, v in ph.items():
                    acc[name] = acc.get(name, 0.0) + v
                stats["collect_ms_sum"] = stats.get("collect_ms_sum", 0.0) + col
                stats["n_collect"] = n_col
                stats["launch_ms"] = sp.launch_ms()
                stats["timed_searches"] = stats.get("timed_searches", 0) + 1
        if n_fail:
            rows = torch.nonzero(fail_flags[:nq], as_tuple=False).squeeze(1)
            sub = PackedSet(q.sign.index_select(0, rows).contiguous(), None, None, int(rows.numel()), q.bits)
            if exact_fallback is None:
                if world > 1:
                    raise RuntimeError("a sharded tensor-core top-K needs exact_fallback")
                redo = topk_exact(sub, d, K, index_base, stripes)
            else:
                redo = exact_fallback(sub)
            if sliced:
                mine = (rows >= rank * per_rank) & (rows < (rank + 1) * per_rank)
                keys.index_copy_(0, rows[mine] - rank * per_rank, redo[mine])
            else:
                keys.index_copy_(0, rows, redo)
        return keys if sliced else keys[:nq]

    # (the dict, not the function itself, is what the closure writes to: a function that names itself in its own body is a
    # reference cycle, and cycles keep the 65 MB key tensors alive until the cyclic collector happens to run)
    finish.verdict = verdict
    return finish if defer else finish()
