"""world_size-2 (and 3) `gloo` runs of the multi-GPU exchange logic in `cmh_b200.sharded`, on CPU, with the numpy
test double of the device passes (`tests/cpu_engine.py`).  Checks N shards == 1 shard: per-query AP, n_rel,
precision@N, PR curve and the merged top-K keys."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _packed_sets(ternary_frac):
    from cmh_b200.engine import PackedSet
    from cmh_b200.synth import EvalShape, make_case
    from oracle import cmh_oracle as orc
    shape = EvalShape("gloo", 9, 301, 20, 24, 0.15, None, (), 77)
    t = make_case(shape, clustered=True, zero_query_frac=0.15, ternary_frac=ternary_frac)

    def mk(codes, labels):
        s, v, nz, _ = orc.pack_codes(codes)
        return PackedSet(torch.from_numpy(s.view(np.int64)), torch.from_numpy(v.view(np.int64)) if ternary_frac else None,
                         torch.from_numpy(orc.pack_labels(labels).view(np.int64)), codes.shape[0], codes.shape[1], labels.shape[1], nz)
    return mk(t["q_img"], t["q_lab"]), mk(t["r_txt"], t["r_lab"]), t


def _worker(rank, world, port, ternary_frac, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cpu_engine
        from cmh_b200 import sharded
        q, d, _ = _packed_sets(ternary_frac)
        lo, hi = sharded.shard_bounds(d.n, world, rank)
        topn = (1, 7, 50, 10_000)
        res = sharded.map_k_sharded(q, d.rows(lo, hi), 40, d.n, topn, eng=cpu_engine, want_pr=True,
                                    ternary=bool(ternary_frac))
        keys = sharded.topk_sharded(q.with_labels(None, 0), d.rows(lo, hi).with_labels(None, 0), 25, lo,
                                    eng=cpu_engine, ternary=bool(ternary_frac))
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), ap=res["ap"].numpy(), n_rel=res["n_rel"].numpy(),
                 map=res["map"].numpy(), prec=res["prec"].numpy(), P=res["pr"][0].numpy(), R=res["pr"][1].numpy(),
                 keys=keys.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,ternary_frac", [(2, 0.0), (2, 0.1), (3, 0.0)])
def test_sharded_equals_single(tmp_path, world, ternary_frac):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ternary_frac, str(tmp_path)), nprocs=world, join=True)
    from oracle import c_oracle, cmh_oracle as orc
    _, _, t = _packed_sets(ternary_frac)
    topn = (1, 7, 50, 10_000)
    want_map, want_ap, want_nrel, want_prec = c_oracle.map_k(t["q_img"], t["r_txt"], t["q_lab"], t["r_lab"], 40, topn)
    wP, wR = orc.pr_curve_counting(t["q_img"], t["r_txt"], t["q_lab"], t["r_lab"])
    want_keys = orc.topk_counting(t["q_img"], t["r_txt"], 25)
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), f"r{r}.npz"))
        assert np.array_equal(z["n_rel"], want_nrel)
        np.testing.assert_allclose(z["ap"], want_ap, rtol=0, atol=1e-12)
        assert abs(float(z["map"][0]) - want_map) < 1e-6
        np.testing.assert_allclose(z["prec"], want_prec, rtol=0, atol=1e-6)
        np.testing.assert_allclose(z["P"], wP, rtol=0, atol=1e-6)
        np.testing.assert_allclose(z["R"], wR, rtol=0, atol=1e-6)
        assert np.array_equal(z["keys"].view(np.uint64), want_keys)       # bit-exact merged ranking


def _comm_worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import cpu_engine
        from cmh_b200 import sharded
        from oracle import cmh_oracle as orc
        comm = sharded.GroupComm(None)
        assert comm.world == world
        # primitives
        t = torch.arange(6, dtype=torch.int64) + 10 * rank
        assert torch.equal(comm.all_reduce_sum(t.clone()), sum(torch.arange(6, dtype=torch.int64) + 10 * r for r in range(world)))
        assert torch.equal(comm.all_reduce_max(t.clone()), torch.arange(6, dtype=torch.int64) + 10 * (world - 1))
        x = (torch.arange(world * 4, dtype=torch.int64).view(world, 2, 2) + 100 * rank)
        got = comm.all_to_all(x)
        for s in range(world):
            assert torch.equal(got[s], torch.arange(world * 4, dtype=torch.int64).view(world, 2, 2)[rank] + 100 * s)
        # the exchange of the tensor-core top-K: global threshold -> partial per-shard lists -> all-to-all by query
        # slice -> merge -> all-gather of the merged slices
        q, d, tt = _packed_sets(0.0)
        K, nq = 25, q.n
        lo, hi = sharded.shard_bounds(d.n, world, rank)
        want = orc.topk_counting(tt["q_img"], tt["r_txt"], K).view(np.int64)
        thr_key = want[:, K - 1] | 0xFFFFFFFF                              # everything up to the K-th distance bucket
        local = orc.topk_counting(tt["q_img"], tt["r_txt"][lo:hi], K).view(np.int64)
        local = np.where(local >= 0, local + lo, local)                    # global row index
        local = np.where((local >= 0) & (local <= thr_key[:, None]), local, -1)  # only rows at or below the threshold
        per_rank = -(-nq // world)
        keys_all = torch.full((per_rank * world, K), -1, dtype=torch.int64)
        keys_all[:nq] = torch.from_numpy(local)
        mine = cpu_engine.topk_merge(comm.all_to_all(keys_all.view(world, per_rank, K)), K)
        keys = comm.all_gather_stack(mine).view(per_rank * world, K)[:nq]
        np.save(os.path.join(out_dir, f"c{rank}.npy"), keys.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_group_comm_exchange(tmp_path, world):
    port = _free_port()
    mp.spawn(_comm_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from oracle import cmh_oracle as orc
    _, _, t = _packed_sets(0.0)
    want = orc.topk_counting(t["q_img"], t["r_txt"], 25)
    for r in range(world):
        assert np.array_equal(np.load(os.path.join(str(tmp_path), f"c{r}.npy")).view(np.uint64), want)


def test_shard_bounds_cover_rows():
    from cmh_b200.sharded import shard_bounds
    for n in (0, 1, 7, 100, 100_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n,world", [(100_000_000, 8), (1_000_003, 3), (5, 4), (0, 2)])
def test_lockstep_stripes_cover_rows_in_order(n, world):
    """Every row belongs to exactly one (rank, stripe) piece; stripe j of every rank lies below stripe j+1 of every
    rank; the stripe descriptions number the local rows by their global index."""
    from cmh_b200 import engine, sharded
    per_rank = [sharded.lockstep_stripes(n, world, r) for r in range(world)]
    n_stripes = len(sharded.LOCKSTEP_FRACTIONS) + 1
    covered = 0
    for j in range(n_stripes):
        pieces = [per_rank[r][0][j] for r in range(world)]
        assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))            # contiguous across the ranks
        if j + 1 < n_stripes:
            assert pieces[-1][1] == per_rank[0][0][j + 1][0]                     # ... and across the stripes
        covered += sum(hi - lo for lo, hi in pieces)
    assert covered == n and per_rank[0][0][0][0] == 0 and per_rank[-1][0][-1][1] == n
    for ranges, stripes in per_rank:
        local = 0
        for (lo, hi), (row, g) in zip(ranges, stripes):
            assert (row, g) == (local, lo)
            local += hi - lo
        norm = engine.check_stripes(stripes, local)
        assert [(a, b, g) for a, b, g in engine.stripe_ranges(norm, local)] == \
               [(r, r + hi - lo, lo) for (lo, hi), (r, _) in zip(ranges, stripes) if hi > lo]
    with pytest.raises(ValueError):
        engine.check_stripes([(1, 0)], 10)


# ---------------------------------------------------------------------------------------------------------------
# the prefix rule across shards: a numpy restatement of what `engine.topk_tc` does between its launches (candidate
# histograms of the rows scanned so far -> exchange -> `tc_choose_kernel`), driven through the real `GroupComm`
# over gloo.  The property under test is the rule's exactness - no row of the true stable top-K is ever excluded,
# ties included - for contiguous shards (all-gather: lower ranks + own prefix at b - 1, everything seen at b) and for
# lockstep stripes (all-reduce, b - 1 on every shard).
# ---------------------------------------------------------------------------------------------------------------
def _choose(hist, need, offset, thr_in):
    """`tc_choose_kernel` (csrc/tc_collect.cu): smallest bucket whose cumulative count reaches `need`, + offset,
    never above thr_in; only buckets <= thr_in are looked at."""
    out = thr_in.copy()
    for q in range(hist.shape[0]):
        cum = 0
        for b in range(0, min(int(thr_in[q]), hist.shape[1] - 1) + 1):
            cum += int(hist[q, b])
            if cum >= need:
                out[q] = min(int(thr_in[q]), b + offset)
                break
    return out


def _prefix_worker(rank, world, port, layout, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cmh_b200 import sharded
        rng = np.random.default_rng(5)
        D, Q, K, bits = 24_000, 24, 40, 16                     # 16-bit codes: every bucket is full of ties
        db = rng.integers(0, 2, (D, bits), dtype=np.int8)
        qs = rng.integers(0, 2, (Q, bits), dtype=np.int8)
        dist_all = (qs[:, None, :] != db[None, :, :]).sum(2).astype(np.int64)          # [Q, D]
        kth = np.sort(dist_all, axis=1, kind="stable")[:, K - 1]
        thr_stat = (kth + 2).astype(np.int64)                   # a loose statistical bound, the same on every shard
        comm = sharded.GroupComm(None)
        fractions = (0.25, 0.6)
        if layout == "lockstep":
            ranges, _ = sharded.lockstep_stripes(D, world, rank, fractions, align=8)
        else:
            lo, hi = sharded.shard_bounds(D, world, rank)
            cuts = [lo] + [lo + int((hi - lo) * f) for f in fractions] + [hi]
            ranges = list(zip(cuts, cuts[1:]))
        nb = bits + 1
        thr = thr_stat.copy()
        cand = [[] for _ in range(Q)]                            # (dist, global row) of the rows kept so far
        for j, (a, b) in enumerate(ranges):
            for q in range(Q):
                rows = np.nonzero(dist_all[q, a:b] <= thr[q])[0] + a
                cand[q] += [(int(dist_all[q, r]), int(r)) for r in rows]
            if j + 1 == len(ranges):
                break
            hist = np.zeros((Q, nb), dtype=np.int64)
            for q in range(Q):
                for dd, _ in cand[q]:
                    hist[q, dd] += 1
            h = torch.from_numpy(hist)
            if layout == "lockstep":
                thr = _choose(comm.all_reduce_sum(h).numpy(), K, -1, thr)
            else:
                every = comm.all_gather_stack(h).numpy()
                thr = _choose(every[:rank + 1].sum(0), K, -1, thr)
                thr = _choose(every.sum(0), K, 0, thr)
        # what the shards kept, merged: must contain the exact stable top-K
        keys = np.full((Q, K), -1, dtype=np.int64)
        for q in range(Q):
            ks = sorted((dd << 33) | r for dd, r in cand[q])[:K]
            keys[q, :len(ks)] = ks
        merged = comm.all_gather_stack(torch.from_numpy(keys)).numpy()                 # [world, Q, K]
        np.savez(os.path.join(out_dir, f"p{rank}.npz"), merged=merged, thr=thr, thr_stat=thr_stat, dist_all=dist_all)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,layout", [(2, "contiguous"), (3, "contiguous"), (2, "lockstep"), (3, "lockstep")])
def test_prefix_rule_across_shards_is_exact(tmp_path, world, layout):
    port = _free_port()
    mp.spawn(_prefix_worker, args=(world, port, layout, str(tmp_path)), nprocs=world, join=True)
    z = [np.load(os.path.join(str(tmp_path), f"p{r}.npz")) for r in range(world)]
    dist_all = z[0]["dist_all"]
    Q, D = dist_all.shape
    K = z[0]["merged"].shape[2]
    order = np.argsort(dist_all, axis=1, kind="stable")[:, :K]
    want = (np.take_along_axis(dist_all, order, 1) << 33) | order
    for r in range(world):
        lists = z[r]["merged"]
        got = np.sort(np.where(lists < 0, np.iinfo(np.int64).max, lists).transpose(1, 0, 2).reshape(Q, -1), axis=1)[:, :K]
        assert np.array_equal(got, want)                                               # nothing of the top K was excluded
        assert np.all(z[r]["thr"] <= z[r]["thr_stat"])
    tightened = [bool(np.any(z[r]["thr"] < z[r]["thr_stat"])) for r in range(world)]
    assert any(tightened)
    if layout == "lockstep":                                                           # every shard, and all alike
        assert all(tightened) and all(np.array_equal(z[r]["thr"], z[0]["thr"]) for r in range(world))
