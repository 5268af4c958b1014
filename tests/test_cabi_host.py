"""CPU-side checks of the boundary: the built library loads and exports every symbol `include/cmh_b200.h`
declares, the ctypes structs mirror the header, argument errors are reported through the C ABI's error
convention, and the Python product refuses to run without a CUDA device (no CPU fallback).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

from cmh_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "cmh_b200.h")


def _header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cmh_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cuda_lib):
    declared = _header_functions()
    assert sorted(_cabi.EXPORTS) == declared, "the _cabi.EXPORTS list and include/cmh_b200.h disagree"
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} is declared in the header but not exported by libcmh_b200.so"
    assert cuda_lib.cmh_abi_version() == _cabi.ABI_VERSION == 2


def test_structs_mirror_header():
    assert ctypes.sizeof(_cabi.CodeSet) == 32            # 3 pointers + int64
    assert ctypes.sizeof(_cabi.Plan) == 12 * 4 + 3 * 8 + 8 + 4 * 4
    src = open(HEADER).read()
    fields = re.search(r"typedef struct cmh_plan \{(.*?)\} cmh_plan;", src, re.S).group(1)
    fields = re.sub(r"/\*.*?\*/", "", fields, flags=re.S)
    names = [n.strip() for decl in fields.split(";") if decl.strip() for n in decl.strip().split(None, 1)[1].split(",")]
    assert names == [f[0] for f in _cabi.Plan._fields_]


def test_search_structs_match_the_library(cuda_lib):
    """The mirrors of `cmh_comm`, `cmh_tc_opts` and `cmh_tc_search` are as long as the library's structs and carry the
    header's field names in the header's order."""
    sizes = (ctypes.c_int32 * 6)()
    assert cuda_lib.cmh_struct_sizes(sizes, 6) == 6
    assert list(sizes) == [ctypes.sizeof(c) for c in (_cabi.CodeSet, _cabi.Plan, _cabi.Comm, _cabi.TcOpts, _cabi.TcSearch)] + [2]
    src = open(HEADER).read()
    for cname, mirror in (("cmh_comm", _cabi.Comm), ("cmh_tc_opts", _cabi.TcOpts), ("cmh_tc_search", _cabi.TcSearch)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), src, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            fp = re.match(r".*\(\*(\w+)\)\(", decl, re.S)        # function pointer member
            if fp:
                names.append(fp.group(1))
                continue
            for n in decl.split(None, 1)[1].split(","):
                names.append(re.sub(r"\[.*?\]|\*", "", n).strip())
        assert names == [f[0] for f in mirror._fields_], cname


def test_search_plan_geometry(cuda_lib):
    """`cmh_tc_search_plan` is host arithmetic on one GPU: launches cut at the pilot stages and the prefix-rule rows,
    every shard of a sharded database would get the same number of exchanges, and a lockstep stripe boundary inside the
    pilot rows is refused (the strict prefix rule would not be valid there)."""
    from cmh_b200 import engine
    p = engine.tc_search_plan(None, 8192, 100_000_000, 100_000_000, 64, 1000, [(0, 0)], 65536)
    assert p.n_stages == 2 and p.n_prefix_cuts == 4 and p.n_spans == 7 and p.exch_width == 1000
    assert [p.span_lo[i] for i in range(1, 7)] == [p.span_hi[i] for i in range(6)] and p.span_hi[6] == 100_000_000
    assert p.n_thr == 7 and p.thr_limit_slot == 2 and p.thr_final_slot == 6
    assert p.seg_total == sum(p.span_n_segs[i] for i in range(7)) and p.workspace_bytes > p.off_cand
    small = engine.tc_search_plan(None, 100, 300_000, 300_000, 128, 50, [(0, 7)], 0, exact_thresholds=True)
    assert small.n_stages == 0 and small.n_prefix_cuts == 0 and small.n_spans == 1 and small.span_index[0] == 7
    assert small.n_sample == 300_000
    short = engine.tc_search_plan(None, 100, 300_000, 300_000, 32, 50, [(0, 0)], 0, exact_thresholds=True)
    assert short.bits == 32                       # codes of up to 32 bits: one K-step per field (NS1)
    assert engine.tc_search_plan(None, 100, 300_000, 300_000, 48, 50, [(0, 0)], 0, exact_thresholds=True).bits == 64
    assert engine.tc_search_plan(None, 100, 300_000, 300_000, 96, 50, [(0, 0)], 0, exact_thresholds=True).bits == 128
    with pytest.raises(ValueError):
        engine.tc_search_plan(None, 100, 1000, 1000, 129, 50, [(0, 0)], 0)           # beyond two packed words: no tensor path
    with pytest.raises(ValueError):
        engine.tc_search_plan(None, 100, 1000, 1000, 64, 5000, [(0, 0)], 0)          # K > 4096


def test_comm_entry_points_reject_bad_arguments(cuda_lib):
    assert cuda_lib.cmh_comm_create_rank(None, 2, 0, None) == -1
    assert cuda_lib.cmh_comm_create(0, None, None) == -1
    assert cuda_lib.cmh_comm_destroy(None) == 0
    assert cuda_lib.cmh_map_k_sharded_workspace_bytes(0, 10, 10, 64, 24, 0, 0) == 0
    assert cuda_lib.cmh_map_k_sharded_workspace_bytes(8, 2100, 25_000, 64, 21, 0, 11) > 0


def test_argument_errors_use_the_error_convention(cuda_lib):
    """Planning is host-only arithmetic: it must reject bad arguments with a negative code and a message, without
    touching a device."""
    plan = _cabi.Plan()
    rc = cuda_lib.cmh_eval_plan(10, 100, 0, 24, 0, 0, ctypes.byref(plan))
    assert rc == -2 and "bits" in _cabi.last_error()
    rc = cuda_lib.cmh_eval_plan(10, 100, 64, 24, 0, 999, ctypes.byref(plan))
    assert rc == -1
    with pytest.raises(ValueError):
        _cabi.check(rc, "cmh_eval_plan")
    rc = cuda_lib.cmh_eval_plan(-1, 100, 64, 24, 0, 0, ctypes.byref(plan))
    assert rc == -1
    assert cuda_lib.cmh_pack_codes(None, 0, 5, 64, 64, None, None, None, None) == -1      # NULL pointers
    assert cuda_lib.cmh_pack_codes(None, 0, 5, 64, 32, None, None, None, None) == -1      # ld < bits
    assert cuda_lib.cmh_topk_merge(None, 0, 1, 1, None, None) == -1


def test_plan_geometry(cuda_lib):
    """The plan is the launch geometry both passes share: chunks cover the database, counters fit 16 bits, and
    the long-code / ternary cases fall back to the warp-per-query design."""
    for nq, nd, bits, nlab, tern in [(2000, 18015, 64, 24, 0), (2100, 193734, 16, 21, 0), (5000, 117218, 128, 80, 0),
                                      (8192, 100_000_000, 64, 0, 0), (7, 700, 2048, 24, 0), (32, 2000, 64, 24, 1),
                                      (1, 1, 1, 1, 0), (0, 0, 64, 24, 0)]:
        plan = _cabi.Plan()
        assert cuda_lib.cmh_eval_plan(nq, nd, bits, nlab, tern, 11, ctypes.byref(plan)) == 0, _cabi.last_error()
        assert plan.nb == (2 * bits + 1 if tern else bits + 1)
        assert plan.chunk_rows % 16 == 0 and plan.chunk_rows <= 65520
        assert plan.n_chunks * plan.chunk_rows >= nd
        assert plan.nq_pad % plan.q_tile == 0 and plan.nq_pad >= nq
        lane = (not tern) and nlab > 0 and 65 <= plan.nb <= 257 and nlab <= 128       # 64 .. 128-bit binary codes with labels
        assert plan.design == (2 if lane else (0 if plan.nb <= 200 else 1))
        assert plan.workspace_bytes > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box WITHOUT a GPU")
def test_product_refuses_to_run_without_cuda():
    from cmh_b200 import calc_utils as cu
    q = torch.ones(2, 16); r = torch.ones(5, 16); l = torch.ones(2, 3); rl = torch.ones(5, 3)
    for call in (lambda: cu.calc_map_k_matrix(q, r, l, rl), lambda: cu.calc_hammingDist(q, r),
                 lambda: cu.calc_neighbor(l, rl), lambda: cu.p_topK(q, r, l, rl, [1]), lambda: cu.pr_curve(q, r, l, rl)):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            call()


def test_product_never_imports_the_oracle():
    """A product path that routes through the oracle voids every parity claim: no module of the package may
    mention it."""
    pkg = os.path.join(ROOT, "clip-based-cross-modal-hashing_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "cmh_oracle" not in src.replace("oracle/cmh_oracle", ""), f"{f} references the oracle"


def test_pack_cache_is_keyed_on_live_tensor_objects():
    from cmh_b200.calc_utils import _PackCache
    c = _PackCache(limit=4)
    t = torch.zeros(4, 8)
    dev = torch.device("cpu")
    c.put(t, "codes", dev, "packed-1")
    assert c.get(t, "codes", dev) == "packed-1"
    t.add_(1)                                   # in-place update bumps the version -> stale entry must miss
    assert c.get(t, "codes", dev) is None
    c.put(t, "codes", dev, "packed-2")
    u = t.clone()
    assert c.get(u, "codes", dev) is None       # a different object never hits, whatever its address
    del t
    assert len(c._entries) == 0                 # entry dies with the tensor


def test_pilot_stages_depend_on_global_sizes_only():
    """Every shard takes part in every threshold refinement: the NUMBER of pilot stages may depend on the database
    and the number of shards, never on a shard's own length."""
    from cmh_b200 import engine
    assert engine.tc_pilot_stages(1_000_000, 1_000_000) == []                     # small database: no pilot
    one = engine.tc_pilot_stages(100_000_000, 100_000_000, 1)
    assert one == [100_000_000 // 512 // 256 * 256, 100_000_000 // 64 // 256 * 256]
    assert engine.tc_pilot_rows(100_000_000) == one[-1]
    for world in (2, 4, 8):                                                       # shards below 64M rows: one stage
        lens = {len(engine.tc_pilot_stages(n, 100_000_000, world)) for n in (0, 5, 12_500_000, 50_000_000)}
        assert lens == {1}
    lens = {len(engine.tc_pilot_stages(n, 1_000_000_000, 8)) for n in (0, 1000, 125_000_000)}
    assert lens == {2}
    assert all(s % 256 == 0 for s in engine.tc_pilot_stages(123_456_789, 1_000_000_000, 8))
