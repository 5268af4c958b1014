"""CPU stand-in for `cmh_b200.engine` - TEST DOUBLE, used only by the gloo tests of the multi-GPU exchange logic
(`cmh_b200.sharded`).  It reproduces the *contract* of the device passes (shard histograms, rank with lower/global
shard totals, partial AP sums, hit counts, top-K keys, merge) with the numpy oracle, so that the host-side
exchange can run with world_size 2 on a box without a GPU.  The product never imports this."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from oracle import cmh_oracle as orc


def _u64(t: Optional[torch.Tensor]) -> Optional[np.ndarray]:
    return None if t is None else t.contiguous().numpy().view(np.uint64)


class RankPass:
    def __init__(self, q, d, *, need_labels=True, max_topn=0, design=-1, ternary=None):
        self.q, self.d, self.need_labels = q, d, need_labels
        self.ternary = (q.valid is not None or d.valid is not None) if ternary is None else bool(ternary)
        self.bits = q.bits
        self.nb = 2 * q.bits + 1 if self.ternary else q.bits + 1
        full = np.uint64((1 << 64) - 1)
        def planes(p):
            s = _u64(p.sign)
            if p.valid is not None:
                return s, _u64(p.valid)
            v = np.full_like(s, full)
            tail = p.bits - 64 * (s.shape[1] - 1)
            if tail < 64:
                v[:, -1] = np.uint64((1 << tail) - 1)
            return s, v
        self.qs, self.qv = planes(q)
        self.ds, self.dv = planes(d)
        self.ql = _u64(q.labels) if need_labels else None
        self.dl = _u64(d.labels) if need_labels else None

    def _bucket(self, i):
        d2 = orc.dist2_packed(self.qs[i], self.qv[i], self.ds, self.dv, self.bits)
        return d2 if self.ternary else d2 // 2

    def _rel(self, i):
        return (self.ql[i][None, :] & self.dl).any(axis=1)

    def hist(self):
        nq = self.q.n
        ha = np.zeros((nq, self.nb), np.int64); hr = np.zeros((nq, self.nb), np.int64)
        for i in range(nq):
            b = self._bucket(i)
            ha[i] = np.bincount(b, minlength=self.nb)
            if self.need_labels:
                hr[i] = np.bincount(b[self._rel(i)], minlength=self.nb)
        self._ha, self._hr = ha, hr
        return (torch.from_numpy(ha.astype(np.int32)), torch.from_numpy(hr.astype(np.int32)) if self.need_labels else None)

    def rank(self, k, topn: Sequence[int] = (), lower=None, glob=None):
        nq, nd = self.q.n, self.d.n
        la = lower[0].numpy().astype(np.int64) if lower is not None else np.zeros_like(self._ha)
        lr = lower[1].numpy().astype(np.int64) if lower is not None else np.zeros_like(self._hr)
        ga = glob[0].numpy().astype(np.int64) if glob is not None else self._ha
        gr = glob[1].numpy().astype(np.int64) if glob is not None else self._hr
        ap_sum = np.zeros(nq, np.float64); n_rel = gr.sum(1)
        hits = np.zeros((nq, len(topn)), np.int32)
        for i in range(nq):
            total = n_rel[i] if k is None else min(int(k), n_rel[i])
            b = self._bucket(i); rel = self._rel(i)
            base_a = np.concatenate(([0], np.cumsum(ga[i])[:-1])) + la[i]
            base_r = np.concatenate(([0], np.cumsum(gr[i])[:-1])) + lr[i]
            run_a = base_a.copy(); run_r = base_r.copy()
            for j in range(nd):
                run_a[b[j]] += 1
                if rel[j]:
                    run_r[b[j]] += 1
                    if run_r[b[j]] <= total:
                        ap_sum[i] += run_r[b[j]] / run_a[b[j]]
                    for t, n in enumerate(topn):
                        if run_a[b[j]] <= n:
                            hits[i, t] += 1
        return (torch.from_numpy(ap_sum), torch.from_numpy(n_rel.astype(np.int64)),
                torch.from_numpy(hits) if len(topn) else None)

    def topk(self, K, index_base=0):
        keys = orc.topk_packed(self.qs, self.qv, self.ds, self.dv, self.bits, K, index_base)
        out = np.full((self.q.n, K), np.uint64((1 << 64) - 1))
        out[:, :keys.shape[1]] = keys
        return torch.from_numpy(out.view(np.int64))


def finalize_map(ap_sum, n_rel, k):
    nr = n_rel.numpy()
    tot = nr if k is None else np.minimum(nr, int(k))
    ap = np.where(tot > 0, ap_sum.numpy() / np.maximum(tot, 1), 0.0)
    return torch.from_numpy(ap), torch.tensor([ap.sum() / max(1, len(ap))], dtype=torch.float32)


def finalize_topn(hits, n_rel, topn, nd_total):
    n = np.minimum(np.asarray(topn, np.float64), nd_total)
    live = (n_rel.numpy() > 0)[:, None]
    p = np.where(live, hits.numpy() / n[None, :], 0.0).sum(0) / max(1, hits.shape[0])
    return torch.from_numpy(p.astype(np.float32))


def finalize_pr(h_all, h_rel, bits, ternary):
    ha = h_all.numpy().astype(np.int64); hr = h_rel.numpy().astype(np.int64)
    step = 2 if ternary else 1
    ca = np.cumsum(ha, 1)[:, ::step][:, :bits + 1].astype(np.float64)
    cr = np.cumsum(hr, 1)[:, ::step][:, :bits + 1].astype(np.float64)
    nr = hr.sum(1).astype(np.float64); live = nr > 0
    P = np.where(live[:, None], cr / np.maximum(ca, 0.1), 0.0)
    R = np.where(live[:, None], cr / np.maximum(nr, 1.0)[:, None], 0.0)
    sup = (P > 0).sum(0).astype(np.float64); sup[sup == 0] = 0.1
    return torch.from_numpy((P.sum(0) / sup).astype(np.float32)), torch.from_numpy((R.sum(0) / sup).astype(np.float32))


def topk_merge(keys_in, K):
    a = keys_in.numpy().view(np.uint64)                       # [G, Q, K]
    cat = np.concatenate(list(a), axis=1)
    return torch.from_numpy(np.sort(cat, axis=1)[:, :K].copy().view(np.int64))
