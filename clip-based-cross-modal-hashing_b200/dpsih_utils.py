"""Drop-in for the reference's ``train/DPSIH/_utils.py`` - the evaluation of SET-valued hash codes (K sub-codes per
item) - with the work done by the sm_100a kernels of ``libcmh_b200.so``.

    from cmh_b200.dpsih_utils import mean_average_precision           # train/DPSIH/_utils.py:4

Reference behaviour reproduced (all `train/DPSIH/_utils.py`):
  * ``qB [Q, K, D]``, ``rB [N, K, D]``: the similarity of two items is the LARGEST of the K x K inner products of their
    sub-codes (:15-20), ``dist = 0.5 * (D - sim)`` (:21) - on packed bits: the SMALLEST of the K x K Hamming distances.
  * relevance ``qL.rL > 0`` (:13); ranking = ascending distance, ties by ascending database index (the reference's
    `torch.argsort` forced stable - the same contract as `calc_map_k_matrix`).
  * ``topk=None -> topk = N`` (:9-10).  The AP is the TEXTBOOK AP@topk: the relevant rows RANKED within the first
    ``topk``, ``mean_j(j / rank_j)`` over those (:22-28); a query without such a row adds 0 and stays in the divisor
    (:24-25, :29).  `calc_map_k_matrix` differs: it averages over the first ``min(k, n_rel)`` relevant rows wherever
    they rank.
  * the result is a 0-d float32 CPU tensor (python ``0.0`` when no query has a hit: the untouched accumulator, :11).
    ``rank`` is unused by the reference; here it only picks the CUDA device for host inputs.
There is no CPU fallback: without a CUDA device or without the built library the call raises.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _cabi
from . import calc_utils as _cu
from . import engine as _e
from ._cabi import Plan, check
from .engine import PackedSet, _ptr, _stream

__all__ = ["mean_average_precision", "set_map_detail"]


def _pack_sets(x, device) -> PackedSet:
    """``[n, K, D]`` +-1 codes -> PackedSet whose row is the K packed sub-codes of an item (``[n, K * words]``)."""
    t = _cu._to_tensor(x)
    if t.dim() != 3:
        raise ValueError(f"set-valued codes must be [n, K, bits], got {tuple(t.shape)}")
    n, K, D = t.shape
    if K < 1 or K > 64:
        raise ValueError(f"{K} sub-codes per item outside [1, 64]")
    flat = _cu.pack_codes(_cu._on(t, device).reshape(n * K, D), device)
    if flat.n_zero:
        raise ValueError("set-valued codes must be +-1 (exact zeros have no packed set form)")
    return PackedSet(flat.sign.view(n, K * flat.words), None, None, n, D)


def set_map_detail(qB, rB, qL, rL, topk: Optional[int] = None, rank=None):
    """Everything the two counting passes produce, on the device: dict(map float32 [1], ap float64 [Q],
    hits int32 [Q] = relevant rows ranked within the first topk, hist_all / hist_rel int32 [Q, D + 1])."""
    dev = _cu._device_for(rank, *(x for x in (qB, rB) if isinstance(x, torch.Tensor)))
    q, d = _pack_sets(qB, dev), _pack_sets(rB, dev)
    if q.bits != d.bits:
        raise RuntimeError(f"code lengths differ: qB has {q.bits} columns, rB has {d.bits}")
    kq, kd = q.sign.shape[1] // q.words, d.sign.shape[1] // d.words
    ql, nlq = _cu.pack_labels(qL, dev)
    dl, nld = _cu.pack_labels(rL, dev)
    if nlq != nld:
        raise RuntimeError(f"label widths differ: {nlq} vs {nld}")
    if ql.shape[0] != q.n or dl.shape[0] != d.n:
        raise RuntimeError(f"labels ({ql.shape[0]}, {dl.shape[0]} rows) do not match codes ({q.n}, {d.n} rows)")
    q, d = q.with_labels(ql, nlq), d.with_labels(dl, nld)
    kk = d.n if topk is None else min(int(topk), d.n)                # `[:topk]` clamps (:22)
    if topk is not None and int(topk) < 0:
        raise ValueError("topk must be None or >= 0")
    nq = q.n
    ap = torch.zeros(nq, dtype=torch.float64, device=dev)
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    hits = torch.zeros((nq, 1), dtype=torch.int32, device=dev)
    res = {"map": out, "ap": ap, "hits": hits[:, 0], "hist_all": None, "hist_rel": None}
    if nq == 0 or d.n == 0 or kk == 0:
        return res
    L = _cabi.lib()
    plan = Plan()
    with torch.cuda.device(dev):
        check(L.cmh_eval_plan_sets(nq, d.n, q.bits, nlq, kq, kd, 1, ctypes.byref(plan)), "cmh_eval_plan_sets")
        ws = torch.empty(max(1, plan.workspace_bytes), dtype=torch.uint8, device=dev)
        h_all = torch.empty((nq, plan.nb), dtype=torch.int32, device=dev)
        h_rel = torch.empty((nq, plan.nb), dtype=torch.int32, device=dev)
        ap_sum = torch.zeros(nq, dtype=torch.float64, device=dev)
        n_rel = torch.zeros(nq, dtype=torch.int64, device=dev)
        qs, ds = q.struct(), d.struct()
        st = _stream(dev)
        check(L.cmh_eval_hist(ctypes.byref(plan), ctypes.byref(qs), ctypes.byref(ds), _ptr(h_all), _ptr(h_rel), _ptr(ws), st),
              "cmh_eval_hist")
        check(L.cmh_eval_rank(ctypes.byref(plan), ctypes.byref(qs), ctypes.byref(ds), kk, None, None, None, None,
                              _cabi.i64_array([kk]), 1, _ptr(hits), _ptr(ap_sum), _ptr(n_rel), _ptr(ws), st), "cmh_eval_rank")
        check(L.cmh_finalize_map_hits(_ptr(ap_sum), _ptr(hits), 1, nq, _ptr(ap), _ptr(out), st), "cmh_finalize_map_hits")
    res.update(hist_all=h_all, hist_rel=h_rel, n_rel=n_rel)
    return res


def mean_average_precision(qB, rB, qL, rL, topk=None, rank=None):
    """`train/DPSIH/_utils.py:4-30`."""
    if _cu._to_tensor(qL).shape[0] == 0:
        raise ZeroDivisionError("float division by zero")            # `mean_AP / num_query` with no query (:29)
    res = set_map_detail(qB, rB, qL, rL, topk, rank)
    m = res["map"].cpu().reshape(())
    return m if float(m) != 0.0 else 0.0
