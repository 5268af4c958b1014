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
    """(rank, world) of ``group`` (None = the default process group when one is initialised; False = this process alone,
    whatever is initialised)."""
    if group is False or not dist.is_available() or not dist.is_initialized():
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


_NATIVE = {}      # process group -> (ctypes pointer to the library's NCCL cmh_comm, rank, world)


def native_comm(group=None):
    """The library's own NCCL transport (`cmh_comm_create_rank`) over the ranks of ``group``: rank 0 draws the NCCL
    unique id, `torch.distributed` ships its 128 bytes, every rank joins on its current CUDA device.  Made once per
    group and kept for the life of the process.  Collective."""
    import ctypes
    from . import _cabi
    key = id(group) if group is not None else None
    if key in _NATIVE:
        return _NATIVE[key][0]
    rank, world = _world(group)
    L = _cabi.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    ident = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        _cabi.check(L.cmh_comm_unique_id(buf), "cmh_comm_unique_id")
        ident.copy_(torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8))
    src = dist.get_global_rank(group, 0) if group is not None else 0
    dist.broadcast(ident, src=src, group=group)
    raw = bytes(ident.cpu().numpy().tobytes())
    out = ctypes.POINTER(_cabi.Comm)()
    _cabi.check(L.cmh_comm_create_rank(raw, world, rank, ctypes.byref(out)), "cmh_comm_create_rank")
    _NATIVE[key] = (out, rank, world)
    return out


class GroupComm:
    """The ranks of a `torch.distributed` process group as the transport of the exchange steps.  On GPUs (NCCL backend)
    `handle()` is the library's own NCCL communicator - the collectives are then issued by the C orchestrator
    (`cmh_topk_tc`, `cmh_map_k_sharded`) in line with its kernels; the tensor methods below serve the Python-level paths
    (`map_k_sharded`, `topk_sharded`) and the gloo tests."""

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = _world(group)

    def handle(self):
        if self.world == 1:
            return None
        if dist.get_backend(self.group) != "nccl":
            raise RuntimeError("the library's transport needs an NCCL process group (one rank per GPU)")
        return native_comm(self.group)

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


def map_k_sharded_native(q, d_shard, k: Optional[int], nd_total: int, topn: Sequence[int] = (), comm=None,
                         want_pr: bool = False, ternary: Optional[bool] = None):
    """`map_k_sharded` as ONE library call (`cmh_map_k_sharded`): the two counting passes and their exchange steps are
    issued by the C ABI over ``comm`` (`GroupComm` -> the library's NCCL transport; None -> one shard).  Same result
    dict."""
    import ctypes
    from . import _cabi
    eng = _engine
    L = _cabi.lib()
    dev = q.device
    tern = (q.valid is not None or d_shard.valid is not None) if ternary is None else bool(ternary)
    if tern:
        if q.valid is None:
            q = eng.PackedSet(q.sign, eng._full_valid(q), q.labels, q.n, q.bits, q.nlab)
        if d_shard.valid is None:
            d_shard = eng.PackedSet(d_shard.sign, eng._full_valid(d_shard), d_shard.labels, d_shard.n, d_shard.bits, d_shard.nlab)
    handle, keep = eng._comm_handle(comm, dev)
    world = 1 if comm is None else int(comm.world)
    topn = [int(t) for t in topn]
    nq, bits = q.n, q.bits
    nbytes = int(L.cmh_map_k_sharded_workspace_bytes(world, nq, d_shard.n, bits, q.nlab, 1 if tern else 0, len(topn)))
    if nbytes == 0:
        raise ValueError(f"cmh_map_k_sharded: cannot plan {nq} x {d_shard.n}, {bits} bits: {_cabi.last_error()}")
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    ap = torch.empty(nq, dtype=torch.float64, device=dev)
    m = torch.zeros(1, dtype=torch.float32, device=dev)
    n_rel = torch.zeros(nq, dtype=torch.int64, device=dev)
    prec = torch.zeros(len(topn), dtype=torch.float32, device=dev) if topn else None
    P = torch.zeros(bits + 1, dtype=torch.float32, device=dev) if want_pr else None
    R = torch.zeros(bits + 1, dtype=torch.float32, device=dev) if want_pr else None
    qs, ds = q.struct(), d_shard.struct()
    with torch.cuda.device(dev):
        rc = L.cmh_map_k_sharded(handle, ctypes.byref(qs), ctypes.byref(ds), bits, q.nlab, 1 if tern else 0,
                                 -1 if k is None else int(k), int(nd_total), _cabi.i64_array(topn), len(topn), eng._ptr(ap),
                                 eng._ptr(m), eng._ptr(n_rel), eng._ptr(prec), eng._ptr(P), eng._ptr(R), eng._ptr(ws), nbytes,
                                 eng._stream(dev))
    if rc and getattr(keep, "error", None) is not None:
        raise keep.error
    _cabi.check(rc, "cmh_map_k_sharded")
    ap.record_stream(torch.cuda.current_stream(dev))
    return {"map": m, "ap": ap, "n_rel": n_rel, "prec": prec, "pr": (P, R) if want_pr else None, "_ws": ws}


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
    """Global top-``K`` keys on the tensor cores (`engine.topk_tc` -> `cmh_topk_tc`): the shards filter with the same
    global thresholds (all-reduced sample and pilot histograms), so each contributes only its share of the ~K rows below
    them; the per-shard lists are exchanged all-to-all by query slice, merged and verified by the slice's rank.  Queries
    that fail the check are redone by `topk_sharded` on every rank.  ``gather=False``: rank r returns only its slice."""
    comm = GroupComm(group)
    stripes = kw.get("stripes")
    return eng.topk_tc(q, d_shard, K, index_base, sample=sample, comm=comm if comm.world > 1 else None, nd_total=nd_total,
                       stats=stats,
                       exact_fallback=lambda sub: topk_sharded(sub, d_shard, K, index_base, group, eng, stripes=stripes),
                       **kw)
