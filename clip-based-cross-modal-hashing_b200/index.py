"""Resident database for large-scale top-K Hamming retrieval (BASELINE.json config 4: 64-bit codes, top-1000,
1M queries x 100M database rows sharded over the GPUs of one box).

The reference only ever ranks inside `calc_map_k_matrix` (`utils/calc_utils.py:30-31`: distance row + full sort per
query); this class is that ranking, truncated to the first K entries, for databases far beyond what a per-query
sort can touch.  The database stays packed in HBM (8 B per 64-bit row: 100M rows = 800 MB); queries stream
through in chunks.  +-1 codes of 64 / 128 bits are searched on the tensor cores (`engine.topk_tc`: int8 GEMM with
a fused candidate filter, global thresholds, all-gather + merge when sharded); anything else - longer codes, small
databases - by two counting passes over the shard (histogram -> threshold -> ordered select).

Exact zeros (`torch.sign(0) == 0`, train/base.py:141) do not take a database off the tensor cores: a one-GPU index
splits its rows once into the +-1 rows (searched on the tensor path in a compacted copy, row numbers mapped back -
compaction keeps the index order, so ties still break by ascending global index) and the few rows holding a zero
(ranked by the ternary counting kernels); the two key lists of a query are merged (`cmh_topk_merge`).  Queries that
hold a zero themselves are ranked by the counting passes against the whole database.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import calc_utils as _cu
from . import engine as _e
from . import sharded as _sh
from .engine import PackedSet


class PendingSearch:
    """Handle of `HammingIndex.search_packed_async`."""

    def __init__(self, finish, sync, keys):
        self._finish, self._sync, self._keys = finish, sync, keys

    def result(self) -> torch.Tensor:
        """int64 [Q, K] keys, usable on the caller's current stream."""
        if self._finish is not None:
            stream, done = self._sync
            with torch.cuda.stream(stream):
                keys = self._finish()                # host sync on the verdict; failed queries redone on that stream
                redo = torch.cuda.Event()
                redo.record(stream)
            cur = torch.cuda.current_stream(keys.device)
            cur.wait_event(redo)
            keys.record_stream(cur)
            self._keys, self._finish = keys, None
        return self._keys


class HammingIndex:
    """Database shard resident on one GPU.

    ``index_base`` is the global row index of the shard's first row; with ``group`` (a `torch.distributed`
    process group, one rank per GPU) `search` returns the global top-K on every rank."""

    # databases at least this long (all shards together) are searched on the tensor cores; shorter ones by the two
    # counting passes, whose fixed costs are lower
    TC_MIN_ROWS = 1_000_000
    SAMPLE_ROWS = 65_536
    _shared_buffers: dict = {}

    def __init__(self, db: PackedSet, index_base: int = 0, group=None, nd_total: Optional[int] = None,
                 sample: Optional[PackedSet] = None, ready=None, stripes=None, assume_binary: bool = False):
        # assume_binary: the caller guarantees +-1 codes of the same length on EVERY shard (`from_packed*`): the
        # constructor then needs no collective and no host sync when nd_total is given
        self._ready = ready                      # [(row_end, event)]: an upload still in flight (`from_packed_host`)
        # a shard made of several row ranges of the database (`sharded.lockstep_stripes`): [(local_row, global_index)]
        self.stripes = _e.check_stripes(stripes, db.n, index_base) if stripes else None
        if db.labels is not None:
            db = db.with_labels(None, 0)
        if db.n and db.sign.data_ptr() % 16:
            # the bulk-copy engine stages 16-byte-aligned tiles; an odd-row view of 64-bit codes is re-homed once
            db = PackedSet(db.sign.clone(), None if db.valid is None else db.valid.clone(), None, db.n, db.bits)
        self.db = db
        self.index_base = int(index_base)
        self.group = group
        # group: a process group; None = the default group when torch.distributed is initialised; False = this process
        # alone (a local index inside a distributed job)
        distributed = group is not False and (group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1))
        if nd_total is None:
            nd_total = db.n
            if distributed:
                t = torch.tensor([db.n], dtype=torch.int64, device=db.device)
                torch.distributed.all_reduce(t, group=group)
                nd_total = int(t.item())
        self.nd_total = int(nd_total)
        self._distributed = bool(distributed)
        # every rank must take the same path: the tensor cores need +-1 codes (no valid plane) on ALL shards
        tc_ok = db.valid is None and _e.tc_supported(db, db)
        if distributed and not assume_binary:
            t = torch.tensor([1 if tc_ok else 0], dtype=torch.int64, device=db.device)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MIN, group=group)
            tc_ok = bool(int(t.item()))
        # plans + candidate scratch are shared by the indices of a device (keyed by geometry, two at most): an index that
        # is rebuilt for every upload of the same shard plans once and reuses the scratch
        self._tc_buffers: dict = HammingIndex._shared_buffers.setdefault(str(db.device), {})
        # a strided sample of the shard: first guess of the per-query thresholds of the tensor-core search
        self.sample = None
        if sample is not None:
            self.sample = sample if tc_ok and self.nd_total >= self.TC_MIN_ROWS else None
        elif tc_ok and self.nd_total >= self.TC_MIN_ROWS:
            # SAMPLE_ROWS is the budget of the whole database: a shard contributes its share (the histograms are summed)
            share = max(4096, self.SAMPLE_ROWS * max(db.n, 1) // max(self.nd_total, 1))
            stride = max(1, db.n // share)
            rows = db.sign[::stride].contiguous()
            self.sample = PackedSet(rows, None, None, rows.shape[0], db.bits)
        # ---- a database with exact zeros on one GPU: +-1 rows on the tensor path, the rest by the ternary counting kernels
        self._hybrid = None
        if (db.valid is not None and not distributed and not self.stripes and self._ready is None
                and self.nd_total >= self.TC_MIN_ROWS and _e.tc_supported(PackedSet(db.sign, None, None, db.n, db.bits),
                                                                          PackedSet(db.sign, None, None, db.n, db.bits))):
            full = _e._full_valid(PackedSet(db.sign[:1], None, None, 1, db.bits))          # [1, words]: all real bits set
            pure_mask = (db.valid == full).all(dim=1)
            pure_rows = torch.nonzero(pure_mask, as_tuple=False).squeeze(1)
            if pure_rows.numel() * 4 >= db.n * 3:                                          # (mostly zeros: not worth a copy)
                mixed_rows = torch.nonzero(~pure_mask, as_tuple=False).squeeze(1)
                pure = PackedSet(db.sign.index_select(0, pure_rows), None, None, int(pure_rows.numel()), db.bits)
                mixed = PackedSet(db.sign.index_select(0, mixed_rows), db.valid.index_select(0, mixed_rows), None,
                                  int(mixed_rows.numel()), db.bits)
                self._hybrid = (HammingIndex(pure, 0, group=False, assume_binary=True), pure_rows + self.index_base,
                                mixed, mixed_rows + self.index_base)

    @classmethod
    def from_codes(cls, rB, device=None, index_base: int = 0, group=None) -> "HammingIndex":
        """Pack float codes ``[D, bits]`` (entries in {-1, 0, +1}) once and keep them on the device."""
        return cls(_cu.pack_codes(rB, device), index_base, group)

    @classmethod
    def from_packed(cls, words: torch.Tensor, bits: int, index_base: int = 0, group=None,
                    nd_total: Optional[int] = None, stripes=None) -> "HammingIndex":
        """Adopt already packed +-1 codes: int64 / uint64-bit-pattern tensor ``[D, ceil(bits/64)]`` on a GPU,
        padding bits zero."""
        if words.dim() != 2 or words.shape[1] != (bits + 63) // 64:
            raise ValueError(f"packed words must be [D, {(bits + 63) // 64}]")
        if not words.is_cuda:
            raise RuntimeError("packed database must be on a CUDA device")
        return cls(PackedSet(words.contiguous().view(torch.int64), None, None, words.shape[0], bits), index_base, group,
                   nd_total, stripes=stripes, assume_binary=nd_total is not None)

    @classmethod
    def from_packed_host(cls, words: torch.Tensor, bits: int, index_base: int = 0, group=None,
                         nd_total: Optional[int] = None, pieces: Optional[int] = None, out: Optional[torch.Tensor] = None,
                         device=None, stripes=None, out_free: Optional["torch.cuda.Event"] = None) -> "HammingIndex":
        """Upload packed +-1 codes from (pinned) host memory WITHOUT waiting for the copy: the rows travel in
        ranges on a copy stream (``pieces`` equal ranges after the pilot rows; None = ranges that follow the search's own
        launches), and the first search scans each range as soon as it has landed (the
        pilot rows first), so the upload of a fresh database hides behind the search that needs it.
        ``out``: optional device tensor [D, words] to upload into (reused between calls).  The upload starts once ``out``
        is no longer read: by default after everything enqueued on the current stream so far; ``out_free`` (an event
        recorded behind the last reader of ``out``) narrows that, so that with two alternating ``out`` buffers the upload
        for the NEXT search runs while the current one is still scanning the other buffer."""
        if words.is_cuda:
            return cls.from_packed(words, bits, index_base, group, nd_total, stripes)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        nwords = (bits + 63) // 64
        if words.dim() != 2 or words.shape[1] != nwords:
            raise ValueError(f"packed words must be [D, {nwords}]")
        n = words.shape[0]
        dst = torch.empty((n, nwords), dtype=torch.int64, device=dev) if out is None else out
        if tuple(dst.shape) != (n, nwords) or dst.dtype != torch.int64 or dst.data_ptr() % 16:
            raise ValueError("out must be a 16-byte aligned int64 [D, words] device tensor")
        words = words.view(torch.int64)
        total = n if nd_total is None else int(nd_total)
        copy_stream = torch.cuda.Stream(dev)
        if out_free is not None:
            copy_stream.wait_event(out_free)                         # the last reader of `out`, named by the caller
        else:
            copy_stream.wait_stream(torch.cuda.current_stream(dev))  # `out` may still be read by earlier work
        world = _sh._world(group)[1]
        stages = _e.tc_pilot_stages(n, total, world)
        n_pilot = stages[-1] if stages else 0
        first = n_pilot if n_pilot else min(n, max(4096, n // 64) // 256 * 256 or n)
        if pieces is None:
            # Upload ranges that follow the search's own launches: the pilot rows, one SHORT range (the first main launch
            # can start ~1 ms into the upload instead of waiting for a third of the database - the link delivers rows
            # about twice as fast as the scan consumes them, so later ranges are never waited for), then the rows up to
            # every cut the search makes anyway (stripe boundaries, or the prefix-rule fractions of one GPU)
            if stripes and len(stripes) > 1:
                later = [int(lo_) for lo_, _ in stripes[1:]]
            else:
                later = [int(n * f) // 256 * 256 for f in (0.3, 0.5, 0.7)]
            early = first + max(256, int(n * 0.065) // 256 * 256)
            cand = [first, early] + later + [n]
        else:
            cand = [first] + [first + (n - first) * (i + 1) // pieces // 256 * 256 for i in range(pieces - 1)] + [n]
        ends = sorted({e for e in cand if 0 < e <= n})
        ready, lo = [], 0
        with torch.cuda.stream(copy_stream):
            t_begin = torch.cuda.Event(enable_timing=True)
            t_begin.record(copy_stream)
            for hi in ends:
                dst[lo:hi].copy_(words[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                ready.append((hi, ev))
                lo = hi
            t_end = torch.cuda.Event(enable_timing=True)
            t_end.record(copy_stream)
        # the threshold sample is a strided gather of the FIRST range once it has landed (no host-side pass over the
        # database): ~SAMPLE_ROWS rows of the whole database, this shard's share of them.  Any subset of the shard's rows
        # is a valid sample - thresholds are only statistical bounds, exactness never depends on them.
        share = max(4096, cls.SAMPLE_ROWS * max(n, 1) // max(total, 1))
        head = ends[0] if ends else 0
        cur = torch.cuda.current_stream(dev)
        if ready:
            cur.wait_event(ready[0][1])
        smp_rows = dst[:head:max(1, head // share)].contiguous() if head else dst[:0]
        sample = PackedSet(smp_rows, None, None, smp_rows.shape[0], bits)
        idx = cls(PackedSet(dst, None, None, n, bits), index_base, group, nd_total, sample=sample, ready=ready,
                  stripes=stripes, assume_binary=nd_total is not None)
        idx.upload_events = (t_begin, t_end)         # (measurement aid: how long the link took for this shard)
        return idx

    def _upload_done(self) -> None:
        if self._ready:
            torch.cuda.current_stream(self.db.device).wait_event(self._ready[-1][1])
            self._ready = None

    def query_slice(self, nq: int):
        """(first query, number of queries) of the slice this rank merges - the rows `search_packed(..., gather=False)`
        returns on this rank."""
        rank, world = _sh._world(self.group)
        per_rank = -(-int(nq) // world)
        lo = min(int(nq), rank * per_rank)
        return lo, max(0, min(per_rank, int(nq) - lo))

    def search_packed(self, q: PackedSet, K: int, stats: Optional[dict] = None, gather: bool = True, defer: bool = False):
        """int64 [Q, K] ascending keys ``(2*dist << 32) | global_row`` (-1 pads rows beyond the database).
        Sharded database, ``gather=False``: the result stays sharded by query slice - rank r returns
        int64 [ceil(Q / world), K], the keys of the queries `query_slice` names (no all-gather of the merged keys).
        ``defer=True``: everything is enqueued on the current stream and a callable comes back; calling it reads the
        verdict (the one host sync of a search), redoes failed queries and returns the keys.  Enqueuing the next chunk
        before resolving the previous one keeps the GPU busy while the host prepares the next call (same stream, same
        scratch: the searches still run one after the other)."""
        if (q.valid is not None and not self._distributed and q.bits == self.db.bits and 1 <= int(K) <= _e.TC_MAX_K
                and (self._hybrid is not None or self.sample is not None)):
            keys = self._split_queries(q, int(K), stats)       # zeros in some queries: the others keep the fast path
            return (lambda: keys) if defer else keys
        if self._hybrid is not None and q.valid is None and q.bits == self.db.bits and 1 <= int(K) <= _e.TC_MAX_K:
            keys = self._search_hybrid(q, int(K))
            return (lambda: keys) if defer else keys
        if self.sample is not None and q.valid is None and q.bits == self.db.bits and 1 <= int(K) <= _e.TC_MAX_K:
            ready, self._ready = self._ready, None       # only the first search can overlap the upload
            return _sh.topk_tc_sharded(q, self.db, int(K), self.index_base, self.nd_total, sample=self.sample,
                                       group=self.group, stats=stats, buffers=self._tc_buffers, ready=ready,
                                       stripes=self.stripes, gather=gather, defer=defer)
        self._upload_done()
        keys = _sh.topk_sharded(q, self.db, int(K), self.index_base, self.group, ternary=None, stripes=self.stripes)
        if not gather:
            rank, world = _sh._world(self.group)
            if world > 1:
                per_rank = -(-q.n // world)
                out = torch.full((per_rank, int(K)), -1, dtype=torch.int64, device=keys.device)
                lo, n = self.query_slice(q.n)
                out[:n] = keys[lo:lo + n]
                keys = out
        return (lambda: keys) if defer else keys

    @staticmethod
    def _remap(keys: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
        """Keys whose low word is a row number of a compacted subset -> the same keys with the global row index."""
        if rows.numel() == 0:
            return keys
        pad = keys < 0
        local = (keys & 0xFFFFFFFF).clamp_(max=rows.numel() - 1)
        return torch.where(pad, keys, (keys & ~0xFFFFFFFF) | rows.index_select(0, local.reshape(-1)).reshape(keys.shape))

    def _split_queries(self, q: PackedSet, K: int, stats) -> torch.Tensor:
        """Queries that hold an exact zero are ranked by the counting passes against the whole database; the +-1 queries
        of the same chunk take the index's fast path (one GPU)."""
        full = _e._full_valid(PackedSet(q.sign[:1], None, None, 1, q.bits))
        is_bin = (q.valid == full).all(dim=1)
        bin_rows = torch.nonzero(is_bin, as_tuple=False).squeeze(1)
        tern_rows = torch.nonzero(~is_bin, as_tuple=False).squeeze(1)
        keys = torch.full((q.n, K), -1, dtype=torch.int64, device=self.db.device)
        if bin_rows.numel():
            q_bin = PackedSet(q.sign.index_select(0, bin_rows), None, None, int(bin_rows.numel()), q.bits)
            keys.index_copy_(0, bin_rows, self.search_packed(q_bin, K, stats))
        if tern_rows.numel():
            q_t = PackedSet(q.sign.index_select(0, tern_rows), q.valid.index_select(0, tern_rows), None,
                            int(tern_rows.numel()), q.bits)
            self._upload_done()
            keys.index_copy_(0, tern_rows, _e.topk_exact(q_t, self.db, K, self.index_base, self.stripes))
        return keys

    def _search_hybrid(self, q: PackedSet, K: int) -> torch.Tensor:
        """+-1 queries against a database that holds exact zeros: tensor path over the +-1 rows, ternary counting kernels
        over the rest, merge."""
        pure_index, pure_rows, mixed, mixed_rows = self._hybrid
        lists = [self._remap(pure_index.search_packed(q, K), pure_rows)]
        if mixed.n:
            # the rows that hold a zero: half-integer distances, keys in the same units (bits - dot)
            lists.append(self._remap(_e.RankPass(q, mixed, need_labels=False, ternary=True).topk(K), mixed_rows))
        return lists[0] if len(lists) == 1 else _e.topk_merge(torch.stack(lists), K)

    def search_packed_async(self, q: PackedSet, K: int, stats: Optional[dict] = None) -> "PendingSearch":
        """`search_packed` without waiting: the search is enqueued on one of two alternating side streams (each with
        its own scratch) and a handle comes back; `handle.result()` reads the verdict and returns the keys on the
        caller's stream.  Issuing chunk i+1 before resolving chunk i keeps two chunks in flight: the small
        kernels, exchanges and ragged last wave of one hide behind the scan of the other.  Handles must be resolved
        in the order they were issued, at most two outstanding."""
        dev = self.db.device
        use_tc = (self.sample is not None and q.valid is None and q.bits == self.db.bits and 1 <= int(K) <= _e.TC_MAX_K)
        if not use_tc:
            return PendingSearch(None, None, self.search_packed(q, K, stats))
        if not hasattr(self, "_lanes"):
            self._lanes = [(torch.cuda.Stream(dev), {}) for _ in range(2)]
            self._turn = 0
        stream, buffers = self._lanes[self._turn % 2]
        self._turn += 1
        ready, self._ready = self._ready, None
        stream.wait_stream(torch.cuda.current_stream(dev))           # the queries (and the database) are ready
        with torch.cuda.stream(stream):
            finish = _sh.topk_tc_sharded(q, self.db, int(K), self.index_base, self.nd_total, sample=self.sample,
                                         group=self.group, stats=stats, buffers=buffers, ready=ready, defer=True,
                                         stripes=self.stripes)
            done = torch.cuda.Event()
            done.record(stream)
        return PendingSearch(finish, (stream, done), None)

    def search(self, qB, K: int):
        """Float query codes in -> (dist float32 [Q, K], index int64 [Q, K]) on the device; pads have index -1."""
        q = _cu.pack_codes(qB, self.db.device)
        if q.bits != self.db.bits:
            raise RuntimeError(f"code lengths differ: queries {q.bits}, database {self.db.bits}")
        keys = self.search_packed(q, K)
        pad = keys < 0
        dist = (keys >> 32).to(torch.float32) * 0.5
        idx = keys & 0xFFFFFFFF
        return dist.masked_fill(pad, float("inf")), idx.masked_fill(pad, -1)
