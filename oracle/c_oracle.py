"""ctypes front end of `oracle/cmh_oracle_c.c` - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The C file is single-threaded; query slices are fanned out over a thread pool here (ctypes drops the GIL), so
full-size parity runs (C2: 4e8 pairs) finish in seconds on the box's host cores.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence, Tuple

import numpy as np

from . import cmh_oracle as _o

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libcmh_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cmh_oracle_c.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _slices(n: int, threads: int):
    threads = max(1, min(threads, n))
    step = -(-n // threads)
    return [(s, min(n, s + step)) for s in range(0, n, step)]


def _threads() -> int:
    return os.cpu_count() or 1


def map_k_packed(qs, qv, ql, ds, dv, dl, bits: int, k: Optional[int] = None,
                 topn: Sequence[int] = ()) -> Tuple[np.ndarray, np.ndarray, Optional[np.ndarray]]:
    """(ap float64 [Q], n_rel int64 [Q], hits int64 [Q, len(topn)] or None) on packed inputs."""
    L = lib()
    nq, nd, words, lwords = qs.shape[0], ds.shape[0], qs.shape[1], ql.shape[1]
    qs, qv, ql, ds, dv, dl = (np.ascontiguousarray(a, dtype=np.uint64) for a in (qs, qv, ql, ds, dv, dl))
    ap = np.zeros(nq, dtype=np.float64)
    nrel = np.zeros(nq, dtype=np.int64)
    tn = np.asarray(list(topn), dtype=np.int64)
    hits = np.zeros((nq, len(tn)), dtype=np.int64) if len(tn) else None

    def run(lo_hi):
        lo, hi = lo_hi
        L.cmh_oracle_map_k(_p(qs[lo:hi]), _p(qv[lo:hi]), _p(ql[lo:hi]), ctypes.c_int64(hi - lo),
                           _p(ds), _p(dv), _p(dl), ctypes.c_int64(nd), words, lwords, int(bits),
                           ctypes.c_int64(-1 if k is None else int(k)), _p(ap[lo:hi]), _p(nrel[lo:hi]),
                           _p(tn) if len(tn) else None, len(tn), _p(hits[lo:hi]) if hits is not None else None)

    with ThreadPoolExecutor(_threads()) as ex:
        list(ex.map(run, _slices(nq, _threads())))
    return ap, nrel, hits


def map_k(qB, rB, query_L, retrieval_L, k: Optional[int] = None, topn: Sequence[int] = ()):
    """Float-code front end: packs with the numpy packer, then `map_k_packed`.
    Returns (mAP float64, ap [Q], n_rel [Q], precision@N float64 [len(topn)] or None)."""
    qs, qv, _, _ = _o.pack_codes(qB)
    ds, dv, _, _ = _o.pack_codes(rB)
    ql, dl = _o.pack_labels(query_L), _o.pack_labels(retrieval_L)
    bits = np.asarray(qB).shape[1]
    ap, nrel, hits = map_k_packed(qs, qv, ql, ds, dv, dl, bits, k, topn)
    prec = None
    if hits is not None:
        n = np.minimum(np.asarray(list(topn), dtype=np.float64), ds.shape[0])
        live = (nrel > 0)[:, None]
        prec = (np.where(live, hits / np.maximum(n, 1.0)[None, :], 0.0)).sum(0) / max(1, qs.shape[0])
    return float(ap.sum() / max(1, qs.shape[0])), ap, nrel, prec


def hist_packed(qs, qv, ql, ds, dv, dl, bits: int) -> Tuple[np.ndarray, np.ndarray]:
    L = lib()
    nq, nd, words = qs.shape[0], ds.shape[0], qs.shape[1]
    lwords = 0 if ql is None else ql.shape[1]
    nb = 2 * bits + 1
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) if a is not None else None for a in (qs, qv, ql, ds, dv, dl)]
    qs, qv, ql, ds, dv, dl = arrs
    h_all = np.zeros((nq, nb), dtype=np.int64)
    h_rel = np.zeros((nq, nb), dtype=np.int64)

    def run(lo_hi):
        lo, hi = lo_hi
        L.cmh_oracle_hist(_p(qs[lo:hi]), _p(qv[lo:hi]), _p(ql[lo:hi]) if ql is not None else None,
                          ctypes.c_int64(hi - lo), _p(ds), _p(dv), _p(dl), ctypes.c_int64(nd),
                          words, lwords, int(bits), _p(h_all[lo:hi]), _p(h_rel[lo:hi]))

    with ThreadPoolExecutor(_threads()) as ex:
        list(ex.map(run, _slices(nq, _threads())))
    return h_all, h_rel


def topk_packed(qs, qv, ds, dv, bits: int, K: int, index_base: int = 0) -> np.ndarray:
    """uint64 [Q, K] ascending keys ``(2*dist << 32) | (index_base + j)``, padded with 2^64-1 when D < K."""
    L = lib()
    nq, nd, words = qs.shape[0], ds.shape[0], qs.shape[1]
    qs, qv, ds, dv = (np.ascontiguousarray(a, dtype=np.uint64) for a in (qs, qv, ds, dv))
    keys = np.empty((nq, int(K)), dtype=np.uint64)

    def run(lo_hi):
        lo, hi = lo_hi
        L.cmh_oracle_topk(_p(qs[lo:hi]), _p(qv[lo:hi]), ctypes.c_int64(hi - lo), _p(ds), _p(dv),
                          ctypes.c_int64(nd), ctypes.c_int64(index_base), words, int(bits), int(K), _p(keys[lo:hi]))

    with ThreadPoolExecutor(_threads()) as ex:
        list(ex.map(run, _slices(nq, _threads())))
    return keys
