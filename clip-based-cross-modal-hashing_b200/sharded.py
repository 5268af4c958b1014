"""Database sharding across the GPUs of one box: one process per GPU (`torch.distributed`, NCCL over NVLink),
queries replicated, database rows split into contiguous index ranges so that "ascending global index" - the tie
order of the stable ranking (`utils/calc_utils.py:31` forced stable) - is (shard, local row).

The path has exactly one exchange step per metric, and nothing else crosses GPUs:

  mAP / precision@N / PR   all-gather of the per-shard bucket histograms ``[Q, nb]`` (all, relevant) after pass 1;
                           every rank derives the global bucket totals and the rows contributed by lower shards,
                           ranks its own rows exactly in pass 2, then one all-reduce(sum) of ``Q`` partial AP sums
                           (+ the precision@N hit counts).  Ranks are integers, so N shards == 1 shard bit for bit;
                           only the order of the final float64 sum differs.
  top-K                    every rank selects its local top-K keys ``(2*dist << 32) | global_row`` (already globally
                           comparable), all-gather, K-way merge kernel.

The reference has no distributed code at all; this is the north star's multi-GPU requirement.
``eng`` is the module providing the device passes (`cmh_b200.engine`); the gloo tests substitute a CPU stand-in to
exercise the exchange logic without a GPU.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import engine as _engine


def shard_bounds(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced row range of ``rank``: the first ``n_rows % world`` shards get one extra row."""
    base, extra = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


LOCKSTEP_FRACTIONS = (0.3, 0.6)   # stripe boundaries of a database sharded in lockstep stripes (= the prefix-rule cuts)


def lockstep_stripes(n_rows: int, world: int, rank: int, fractions=LOCKSTEP_FRACTIONS, align: int = 256):
    """Split a database of ``n_rows`` into ``len(fractions) + 1`` global stripes and every stripe into ``world``
    contiguous pieces: rank r holds piece r of every stripe.  All shards then walk the database front to back in
    lockstep - after stripe j everything any shard has scanned has a lower index than everything still to come - so the
    exact prefix rule of the tensor-core top-K tightens as early on 8 GPUs as on one (with one contiguous range per
    shard, shard 0 never sees a lower-index row outside its own prefix).
    Returns ``(ranges, stripes)``: the global row ranges ``[(lo, hi), ...]`` this rank holds, in local order, and the
    ``[(local_row, global_index), ...]`` description `engine.topk_tc` / `HammingIndex` take."""
    unit = max(1, int(align)) * int(world)
    edges = [0] + [min(int(n_rows), int(n_rows * f) // unit * unit) for f in fractions] + [int(n_rows)]
    ranges, stripes, local = [], [], 0
    for a, b in zip(edges, edges[1:]):
        lo, hi = shard_bounds(b - a, world, rank)
        ranges.append((a + lo, a + hi))
        stripes.append((local, a + lo))
        local += hi - lo
    return ranges, stripes


def _world(group) -> Tuple[int, int]:
    if not dist.is_available() or not dist.is_initialized():
        return 0, 1
    return dist.get_rank(group), dist.get_world_size(group)


def _all_gather_stack(x: torch.Tensor, world: int, group) -> torch.Tensor:
    """[world, *x.shape] from every rank's ``x`` (concatenation along dim 0 - the layout both NCCL and gloo accept)."""
    x = x.contiguous()
    flat = torch.empty((world * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(flat, x, group=group)
    return flat.view((world,) + tuple(x.shape))


def exchange_histograms(h_all: torch.Tensor, h_rel: Optional[torch.Tensor], group=None):
    """All-gather the shard histograms and reduce them to what pass 2 needs.

    Returns ((lower_all, lower_rel), (global_all, global_rel)) as int32 tensors with uint32 bit patterns
    (sums are taken in int64 and truncated, i.e. modulo 2^32 like the kernels' own counters)."""
    rank, world = _world(group)
    both = torch.stack([h_all, h_rel if h_rel is not None else torch.zeros_like(h_all)])  # [2, Q, nb]
    if world == 1:
        zero = torch.zeros_like(both)
        return (zero[0], zero[1]), (both[0], both[1])
    gathered = _all_gather_stack(both, world, group)
    wide = gathered.to(torch.int64) & 0xFFFFFFFF
    lower = wide[:rank].sum(0).to(torch.int32) if rank > 0 else torch.zeros_like(both)
    glob = wide.sum(0).to(torch.int32)
    return (lower[0].contiguous(), lower[1].contiguous()), (glob[0].contiguous(), glob[1].contiguous())


def map_k_sharded(q, d_shard, k: Optional[int], nd_total: int, topn: Sequence[int] = (), group=None,
                  eng=_engine, want_pr: bool = False, ternary: Optional[bool] = None):
    """One direction of `calc_map_k_matrix` with the database sharded over the ranks of ``group``.

    q        packed queries (replicated on every rank, labels attached)
    d_shard  this rank's contiguous database rows (labels attached)
    ternary  must be the same on every rank: True when ANY shard (or the queries) holds exact-zero entries
    Returns dict(map float32 [1], ap float64 [Q], n_rel int64 [Q], prec float32 [len(topn)] | None,
                 pr (P, R) | None) - identical on every rank."""
    rp = eng.RankPass(q, d_shard, need_labels=True, max_topn=len(topn), ternary=ternary)
    h_all, h_rel = rp.hist()
    lower, glob = exchange_histograms(h_all, h_rel, group)
    ap_sum, n_rel, hits = rp.rank(k, topn, lower=lower, glob=glob)
    _, world = _world(group)
    if world > 1:
        dist.all_reduce(ap_sum, op=dist.ReduceOp.SUM, group=group)
        if hits is not None:
            dist.all_reduce(hits, op=dist.ReduceOp.SUM, group=group)
    ap, m = eng.finalize_map(ap_sum, n_rel, k)
    prec = eng.finalize_topn(hits, n_rel, topn, nd_total) if len(topn) else None
    pr = eng.finalize_pr(glob[0], glob[1], q.bits, rp.ternary) if want_pr else None
    return {"map": m, "ap": ap, "n_rel": n_rel, "prec": prec, "pr": pr}


class GroupComm:
    """The exchange steps of `engine.topk_tc` over the ranks of a process group (NCCL on GPUs)."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = _world(group)

    def all_reduce_sum(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_max(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def all_gather_stack(self, t: torch.Tensor) -> torch.Tensor:
        if self.world == 1:
            return t.unsqueeze(0)
        return _all_gather_stack(t, self.world, self.group)

    def all_to_all(self, t: torch.Tensor) -> torch.Tensor:
        """t [world, ...]: slice s goes to rank s; returns [world, ...] with slice s received from rank s."""
        if self.world == 1:
            return t
        t = t.contiguous()
        if dist.get_backend(self.group) == "gloo":       # gloo has no all-to-all: gather everything, keep my column
            rank = dist.get_rank(self.group)
            return _all_gather_stack(t, self.world, self.group)[:, rank].contiguous()
        out = torch.empty_like(t)
        dist.all_to_all_single(out, t, group=self.group)
        return out


def topk_sharded(q, d_shard, K: int, index_base: int, group=None, eng=_engine,
                 ternary: Optional[bool] = None, stripes=None) -> torch.Tensor:
    """Global top-``K`` keys int64 [Q, K] (ascending, ``-1`` pads) - identical on every rank.
    ``index_base`` is the global index of this shard's first row (``stripes``: `engine.check_stripes`, a shard made
    of several row ranges).  The exact two-pass (popc) path: every rank selects its local top-K, one all-gather,
    K-way merge."""
    if stripes and len(stripes) > 1:
        local = eng.topk_exact(q, d_shard, K, index_base, stripes, ternary=ternary)
    else:
        rp = eng.RankPass(q, d_shard, need_labels=False, ternary=ternary)
        local = rp.topk(K, index_base)
    _, world = _world(group)
    if world == 1:
        return local
    return eng.topk_merge(_all_gather_stack(local, world, group), K)


def topk_tc_sharded(q, d_shard, K: int, index_base: int, nd_total: int, sample=None, group=None, eng=_engine,
                    stats: Optional[dict] = None, **kw) -> torch.Tensor:
    """Global top-``K`` keys on the tensor cores (`engine.topk_tc`): the shards filter with the same global
    thresholds (all-reduced sample and pilot histograms, a few KB per query chunk), so each contributes only its
    share of the ~K rows below them; the per-shard results are all-gathered and merged, and the merged K-th key is
    verified against the thresholds.  Queries that fail the check are redone by `topk_sharded` on every rank."""
    comm = GroupComm(group)
    stripes = kw.get("stripes")
    return eng.topk_tc(q, d_shard, K, index_base, sample=sample, comm=comm, nd_total=nd_total, stats=stats,
                       exact_fallback=lambda sub: topk_sharded(sub, d_shard, K, index_base, group, eng, stripes=stripes),
                       **kw)
