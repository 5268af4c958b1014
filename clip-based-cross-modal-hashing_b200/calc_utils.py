"""Drop-in for the reference's ``utils/calc_utils.py`` - same names, same argument meaning, same return types -
with the work done by the sm_100a kernels of ``libcmh_b200.so``.

    from cmh_b200.calc_utils import calc_map_k_matrix as calc_map_k     # train/base.py:11
    from cmh_b200.calc_utils import calc_neighbor                        # train/MITH/hash_train.py:12

Reference behaviour reproduced (all `utils/calc_utils.py`):
  * ``dist = 0.5 * (bits - q.r)`` (:8-13); exact zeros in a code (``torch.sign(0)``) give half-integer distances.
  * relevance ``qL.rL > 0`` (:26); queries without a relevant row are skipped but stay in the divisor (:28-29,38).
  * ranking = ascending distance, ties by ascending database index (the reference's `torch.sort` forced stable).
  * ``k=None -> k = D`` (:23-24); ``total = min(k, n_rel)``; AP over the first ``total`` relevant rows wherever
    they rank (:34-37) - not the textbook mAP@k.
  * the result is a 0-d float32 CPU tensor; ``rank`` is only used to pick the CUDA device for host inputs
    (the reference ignores it).
There is no CPU fallback: without a CUDA device or without the built library every function raises.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import engine as _e
from .engine import PackedSet

__all__ = ["calc_hammingDist", "calc_map_k_matrix", "calc_map_k", "calc_neighbor", "p_topK", "pr_curve",
           "topk_hamming", "map_k_detail", "pack_codes", "pack_labels", "clear_cache", "set_cache"]


# ---------------------------------------------------------------------------------------------------------------
# input handling
# ---------------------------------------------------------------------------------------------------------------
def _device_for(rank, *tensors) -> torch.device:
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("cmh_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    if isinstance(rank, torch.device):
        return rank
    if isinstance(rank, str):
        return torch.device(rank)
    return torch.device("cuda", int(rank) if rank is not None else torch.cuda.current_device())


def _to_tensor(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.detach()
    return torch.as_tensor(x)


def _on(x: torch.Tensor, device: torch.device) -> torch.Tensor:
    if x.device == device:
        return x
    return x.to(device)               # the reference moves everything to one place (:19-21); so do we


# Packed forms are cached per *tensor object* (weak reference + in-place version counter), never per data
# pointer: `valid()` makes four calls on the same four buffers (train/base.py:259-262) and every buffer is packed
# once.  A dead tensor or one modified through torch can never hit.  Writes that bypass torch's version counter
# (`t.data[...] = ...`, a `torch.from_numpy` buffer changed through numpy, custom kernels) are invisible to it: call
# `clear_cache()` after such a write, or switch the cache off with `set_cache(False)` (the reference recomputes from
# the live buffer on every call).
class _PackCache:
    def __init__(self, limit: int = 16):
        self._entries: Dict[Tuple[int, str, str], Tuple[weakref.ref, int, object]] = {}
        self._limit = limit

    def get(self, t: torch.Tensor, kind: str, device: torch.device):
        key = (id(t), kind, str(device))
        ent = self._entries.get(key)
        if ent is None:
            return None
        ref, version, value = ent
        if ref() is t and t._version == version:
            return value
        del self._entries[key]
        return None

    def put(self, t: torch.Tensor, kind: str, device: torch.device, value) -> None:
        if len(self._entries) >= self._limit:
            self._entries.pop(next(iter(self._entries)))
        key = (id(t), kind, str(device))
        try:
            ref = weakref.ref(t, lambda _r, k=key, s=self: s._entries.pop(k, None))
        except TypeError:
            return
        self._entries[key] = (ref, t._version, value)

    def clear(self) -> None:
        self._entries.clear()


_cache = _PackCache()
_cache_enabled = True


def clear_cache() -> None:
    _cache.clear()


def set_cache(enabled: bool) -> None:
    """Enable / disable the pack cache (disabled: every call packs its inputs afresh, like the reference)."""
    global _cache_enabled
    _cache_enabled = bool(enabled)
    if not _cache_enabled:
        _cache.clear()


class _PendingCodes:
    """Codes whose pack kernel is enqueued but whose counters (zeros / values outside {-1, 0, +1}) have not been read."""

    def __init__(self, owner, kind, device, sign, valid, n, bits, counters, binarize):
        self.owner, self.kind, self.device = owner, kind, device
        self.sign, self.valid, self.n, self.bits, self.counters, self.binarize = sign, valid, n, bits, counters, binarize

    def resolve(self, n_zero: int, n_odd: int) -> PackedSet:
        if n_odd and not self.binarize:
            raise ValueError(f"{n_odd} code entries are outside {{-1, 0, +1}}: pass sign()-binarised codes "
                             "(train/base.py:141) - the reference's float dot product on raw activations is not "
                             "a Hamming distance")
        ps = PackedSet(self.sign, self.valid if n_zero else None, None, self.n, self.bits, 0, n_zero)
        if self.owner is not None:
            _cache.put(self.owner, self.kind, self.device, ps)
        return ps


class _PendingLabels:
    def __init__(self, owner, device, masks, nlab, counters):
        self.owner, self.device, self.masks, self.nlab, self.counters = owner, device, masks, nlab, counters

    def resolve(self, n_neg: int, _unused: int = 0):
        if n_neg:
            raise ValueError("labels must be non-negative multi-hot (the reference's `dot > 0` relevance, "
                             "utils/calc_utils.py:26, is a set intersection only then)")
        out = (self.masks, self.nlab)
        if self.owner is not None:
            _cache.put(self.owner, "labels", self.device, out)
        return out


def _pack_codes_enqueue(x, device, binarize: bool, counters: torch.Tensor):
    """Enqueue K1 on codes; returns a finished PackedSet (already packed input / cache hit) or a `_PendingCodes` whose
    two counters live in ``counters`` (int64 [2], zeroed by the caller)."""
    if isinstance(x, PackedSet):                     # already packed (engine level)
        return x
    if hasattr(x, "packed") and hasattr(x, "put"):   # codes.CodeBuffer: binarised and packed at the source
        return x.packed()
    t = _to_tensor(x)
    owner = x if isinstance(x, torch.Tensor) and _cache_enabled else None
    kind = "codes-b" if binarize else "codes"
    if owner is not None:
        hit = _cache.get(owner, kind, device)
        if hit is not None:
            return hit
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"codes must be [n, bits], got {tuple(t.shape)}")
    # pinned host codes are read by the pack kernel straight over the link (no staging copy of the floats in HBM)
    td = t if (not t.is_cuda and t.is_pinned() and t.stride(-1) == 1) else _on(t, device)
    sign, valid = _e.pack_codes_device(td, counters, device)
    return _PendingCodes(owner, kind, device, sign, valid, td.shape[0], td.shape[1], counters, binarize)


def _pack_labels_enqueue(L, device, counters: torch.Tensor):
    t = _to_tensor(L)
    owner = L if isinstance(L, torch.Tensor) and _cache_enabled else None
    if owner is not None:
        hit = _cache.get(owner, "labels", device)
        if hit is not None:
            return hit
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError(f"labels must be [n, nlab], got {tuple(t.shape)}")
    td = _on(t, device)
    masks = _e.pack_labels_device(td, counters)
    return _PendingLabels(owner, device, masks, td.shape[1], counters)


def _resolve(items, counters: torch.Tensor):
    """ONE device-to-host read for everything that was enqueued: the counters decide whether a valid plane is needed
    (exact zeros) and whether the inputs were legal."""
    pending = [it for it in items if isinstance(it, (_PendingCodes, _PendingLabels))]
    if not pending:
        return items
    host = counters.tolist()                         # the one sync of the pack stage
    out = []
    for i, it in enumerate(items):
        if isinstance(it, (_PendingCodes, _PendingLabels)):
            out.append(it.resolve(int(host[2 * i]), int(host[2 * i + 1])))
        else:
            out.append(it)
    return out


def pack_codes(x, device=None, *, binarize: bool = False) -> PackedSet:
    """Sign + bit-pack ``[n, bits]`` codes on the device (kernel K1).  Entries must be in {-1, 0, +1} (what
    `torch.sign` / the argmax heads emit, train/base.py:141-158) unless ``binarize`` is set, in which case the
    sign of any real value is taken (the duplicate API `utils/utils.py:77-78` does that)."""
    if isinstance(x, PackedSet) or (hasattr(x, "packed") and hasattr(x, "put")):
        return _pack_codes_enqueue(x, device, binarize, None)
    device = _device_for(device, _to_tensor(x))
    counters = torch.zeros(2, dtype=torch.int64, device=device)
    return _resolve([_pack_codes_enqueue(x, device, binarize, counters)], counters)[0]


def pack_labels(L, device=None) -> Tuple[torch.Tensor, int]:
    """Multi-hot labels ``[n, nlab]`` -> (int64 masks [n, ceil(nlab/64)] on the device, nlab)."""
    device = _device_for(device, _to_tensor(L))
    counters = torch.zeros(2, dtype=torch.int64, device=device)
    return _resolve([_pack_labels_enqueue(L, device, counters)], counters)[0]


def _prepare(qB, rB, query_L, retrieval_L, rank, binarize=False) -> Tuple[PackedSet, PackedSet]:
    dev = None
    for x in (qB, rB):                               # packed inputs (PackedSet, CodeBuffer) decide the device
        if not isinstance(x, torch.Tensor) and isinstance(getattr(x, "device", None), torch.device):
            dev = x.device
    if dev is None:
        dev = _device_for(rank, *(x for x in (qB, rB) if isinstance(x, torch.Tensor)))
    # all four pack kernels are enqueued before their counters are read: one host sync instead of four
    counters = torch.zeros(8, dtype=torch.int64, device=dev)
    items = [_pack_codes_enqueue(qB, dev, binarize, counters[0:2]), _pack_codes_enqueue(rB, dev, binarize, counters[2:4])]
    if query_L is not None:
        items += [_pack_labels_enqueue(query_L, dev, counters[4:6]), _pack_labels_enqueue(retrieval_L, dev, counters[6:8])]
    items = _resolve(items, counters)
    q, d = items[0], items[1]
    if q.bits != d.bits:
        raise RuntimeError(f"code lengths differ: qB has {q.bits} columns, rB has {d.bits}")   # torch.mm would raise
    if query_L is not None:
        (ql, nlq), (dl, nld) = items[2], items[3]
        if nlq != nld:
            raise RuntimeError(f"label widths differ: {nlq} vs {nld}")
        if ql.shape[0] != q.n or dl.shape[0] != d.n:
            raise RuntimeError(f"labels ({ql.shape[0]}, {dl.shape[0]} rows) do not match codes ({q.n}, {d.n} rows)")
        q, d = q.with_labels(ql, nlq), d.with_labels(dl, nld)
    return q, d


# ---------------------------------------------------------------------------------------------------------------
# the reference's functions
# ---------------------------------------------------------------------------------------------------------------
def calc_hammingDist(B1, B2):
    """`utils/calc_utils.py:8-13`: ``0.5 * (B2.shape[1] - B1 @ B2.T)``, float32 ``[m, n]``; a 1-D ``B1`` is one
    row.  The result lives where ``B1`` lives (host inputs get a host result, like the reference)."""
    b1, b2 = _to_tensor(B1), _to_tensor(B2)
    dev = _device_for(None, b1, b2)
    q, d = pack_codes(B1, dev), pack_codes(B2, dev)
    if q.bits != d.bits:
        raise RuntimeError(f"size mismatch: B1 has {q.bits} columns, B2 has {d.bits}")
    out = _e.hamming_dense(q, d)
    return out if b1.is_cuda else out.cpu()


def calc_neighbor(label1, label2):
    """`utils/calc_utils.py:42-45`: float32 ``[m, n]``, 1.0 where the rows share a label."""
    a, b = _to_tensor(label1), _to_tensor(label2)
    dev = _device_for(None, a, b)
    ma, na = pack_labels(label1, dev)
    mb, nb = pack_labels(label2, dev)
    if na != nb:
        raise RuntimeError(f"size mismatch: {na} vs {nb} label columns")
    out = _e.neighbor_dense(ma, mb, na)
    return out if a.is_cuda else out.cpu()


def map_k_detail(qB, rB, query_L, retrieval_L, k=None, rank=0, topn: Sequence[int] = (), design: int = -1,
                 binarize: bool = False):
    """Everything the two counting passes produce for one direction, still on the device:
    dict(map float32 [1], ap float64 [Q], n_rel int64 [Q], prec float32 [len(topn)] or None,
         hist_all / hist_rel int32 [Q, nb], ternary bool, bits int)."""
    q, d = _prepare(qB, rB, query_L, retrieval_L, rank, binarize)
    if k is not None and int(k) < 0:
        raise ValueError("k must be None or >= 0")
    rp = _e.RankPass(q, d, need_labels=True, max_topn=len(topn), design=design)
    h_all, h_rel = rp.hist()
    ap_sum, n_rel, hits = rp.rank(k, topn)
    ap, m = _e.finalize_map(ap_sum, n_rel, k)
    prec = _e.finalize_topn(hits, n_rel, topn, d.n) if len(topn) else None
    return {"map": m, "ap": ap, "n_rel": n_rel, "prec": prec, "hist_all": h_all, "hist_rel": h_rel,
            "ternary": rp.ternary, "bits": q.bits, "hits": hits}


def calc_map_k_matrix(qB, rB, query_L, retrieval_L, k=None, rank=0):
    """`utils/calc_utils.py:16-39`.  Returns the mAP as a 0-d float32 CPU tensor (python ``0.0`` when there is no
    query, as the reference's untouched accumulator would be)."""
    res = map_k_detail(qB, rB, query_L, retrieval_L, k, rank)
    if res["ap"].shape[0] == 0:
        return 0.0
    return res["map"].cpu().reshape(())


calc_map_k = calc_map_k_matrix


# ---------------------------------------------------------------------------------------------------------------
# north-star additions (not in the reference; definitions frozen in oracle/cmh_oracle.py)
# ---------------------------------------------------------------------------------------------------------------
def p_topK(qB, rB, query_L, retrieval_L, K: Sequence[int], rank=0):
    """precision@N for every N in ``K``: mean over queries (label-free queries contribute 0) of the share of
    relevant rows among the first ``min(N, D)`` of the stable ranking.  float32 CPU tensor ``[len(K)]``."""
    K = [int(v) for v in K]
    out = torch.zeros(len(K), dtype=torch.float32)
    for lo in range(0, len(K), 64):                      # CMH_MAX_TOPN cutoffs per pass
        part = K[lo:lo + 64]
        res = map_k_detail(qB, rB, query_L, retrieval_L, None, rank, topn=part)
        out[lo:lo + len(part)] = res["prec"].cpu()
    return out


def pr_curve(qB, rB, query_L, retrieval_L, rank=0):
    """Hamming-radius precision / recall curve, radii 0..bits.  Two float32 CPU tensors ``[bits + 1]``."""
    q, d = _prepare(qB, rB, query_L, retrieval_L, rank)
    rp = _e.RankPass(q, d, need_labels=True)
    h_all, h_rel = rp.hist()
    P, R = _e.finalize_pr(h_all, h_rel, q.bits, rp.ternary)
    return P.cpu(), R.cpu()


_TC_MIN_ROWS = 1_000_000


def topk_hamming(qB, rB, K: int, rank=0):
    """First ``K`` entries of the stable ascending-distance ranking of every query (`utils/calc_utils.py:30-31`
    truncated).  Returns (dist float32 [Q, K'], index int64 [Q, K']) on the device, K' = min(K, D)."""
    q, d = _prepare(qB, rB, None, None, rank)
    kk = min(int(K), d.n)
    if kk <= 0 or q.n == 0:
        return (torch.empty((q.n, 0), dtype=torch.float32, device=q.device),
                torch.empty((q.n, 0), dtype=torch.int64, device=q.device))
    if d.n >= _TC_MIN_ROWS and _e.tc_supported(q, d, kk):
        # +-1 codes of 64 / 128 bits against a large database: int8 tcgen05 GEMM with the fused candidate filter
        stride = max(1, d.n // 65_536)
        rows = d.sign[::stride].contiguous()
        keys = _e.topk_tc(q, d, kk, sample=_e.PackedSet(rows, None, None, rows.shape[0], d.bits))
    elif d.n >= _TC_MIN_ROWS and bool(_e._cabi.lib().cmh_tc_supported(d.bits, 0)) and kk <= _e.TC_MAX_K:
        # exact zeros in a large database (or in some queries): the +-1 rows / queries still go to the tensor path
        # (`index.HammingIndex`: hybrid split of the rows, split of the queries)
        from .index import HammingIndex
        keys = HammingIndex(d, 0, group=False).search_packed(q, kk)
    else:
        keys = _e.RankPass(q, d, need_labels=False).topk(kk)
    return (keys >> 32).to(torch.float32) * 0.5, keys & 0xFFFFFFFF
