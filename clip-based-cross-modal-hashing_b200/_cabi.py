"""ctypes binding of ``libcmh_b200.so`` (declared in ``include/cmh_b200.h``).

This is the only place the Python host code touches the C ABI.  There is deliberately no fallback: when the
library is missing or fails to load, every entry point raises - the oracle under ``oracle/`` is test
infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading
from typing import List, Optional

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_PKG_DIR, "csrc")
LIB_PATH = os.path.join(_PKG_DIR, "libcmh_b200.so")
INCLUDE_DIR = os.path.join(os.path.dirname(_PKG_DIR), "include")
SOURCES = ("api.cu", "pack.cu", "dense.cu", "eval_tile.cu", "eval_warp.cu", "eval_lane.cu", "eval_host.cu", "peaks.cu", "tc_collect.cu",
           "tc_search.cu", "comm.cu", "sharded.cu", "head.cu")
NVCC_FLAGS = ("-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-t", "0", "-ldl")

CMH_MAX_TOPN = 64
CMH_MAX_BITS = 4096
DTYPE_CODES = {"float32": 0, "float16": 1, "bfloat16": 2, "float64": 3, "int8": 4, "int32": 5, "int64": 6, "uint8": 7}

# every symbol include/cmh_b200.h declares (tests check that the built library exports each of them)
EXPORTS = (
    "cmh_abi_version", "cmh_last_error", "cmh_device_info", "cmh_launch_count", "cmh_measure_popc_peak",
    "cmh_pack_codes", "cmh_pack_scatter", "cmh_hash_head_pack", "cmh_unpack_codes", "cmh_pack_labels", "cmh_synth_codes",
    "cmh_hamming_dense", "cmh_neighbor_dense",
    "cmh_eval_plan", "cmh_eval_plan_design", "cmh_eval_plan_sets", "cmh_eval_hist", "cmh_eval_rank", "cmh_finalize_map_hits",
    "cmh_finalize_map", "cmh_finalize_topn", "cmh_finalize_pr_workspace_bytes", "cmh_finalize_pr",
    "cmh_map_k_workspace_bytes", "cmh_map_k",
    "cmh_topk", "cmh_topk_merge",
    "cmh_tc_supported", "cmh_tc_plan", "cmh_tc_collect", "cmh_tc_probe", "cmh_tc_set_workers", "cmh_tc_cand_hist", "cmh_tc_choose", "cmh_tc_choose_prefix", "cmh_tc_choose_seen",
    "cmh_topk_threshold", "cmh_topk_finalize", "cmh_topk_verify", "cmh_topk_merge_verify",
    "cmh_comm_create", "cmh_comm_unique_id", "cmh_comm_create_rank", "cmh_comm_create_loopback", "cmh_comm_destroy",
    "cmh_topk_sharded", "cmh_map_k_sharded_workspace_bytes", "cmh_map_k_sharded",
    "cmh_tc_default_opts", "cmh_tc_pilot_stages", "cmh_struct_sizes", "cmh_tc_search_plan",
    "cmh_tc_timing_create", "cmh_tc_timing_destroy", "cmh_tc_timing_read", "cmh_tc_timing_launches", "cmh_topk_tc",
)
ABI_VERSION = 2


class CodeSet(ctypes.Structure):
    """``cmh_codeset``"""
    _fields_ = [("sign", ctypes.c_void_p), ("valid", ctypes.c_void_p), ("labels", ctypes.c_void_p),
                ("n", ctypes.c_int64)]


class Plan(ctypes.Structure):
    """``cmh_plan``"""
    _fields_ = [("bits", ctypes.c_int32), ("words", ctypes.c_int32), ("nlab", ctypes.c_int32),
                ("lwords", ctypes.c_int32), ("ternary", ctypes.c_int32), ("max_topn", ctypes.c_int32),
                ("nb", ctypes.c_int32), ("design", ctypes.c_int32), ("q_tile", ctypes.c_int32),
                ("n_qtiles", ctypes.c_int32), ("chunk_rows", ctypes.c_int32), ("n_chunks", ctypes.c_int32),
                ("nq", ctypes.c_int64), ("nd", ctypes.c_int64), ("nq_pad", ctypes.c_int64),
                ("workspace_bytes", ctypes.c_uint64), ("kq", ctypes.c_int32), ("kd", ctypes.c_int32),
                ("ap_mode", ctypes.c_int32), ("reserved", ctypes.c_int32)]


TC_MAX_STRIPES, TC_MAX_STAGES, TC_MAX_CUTS, TC_MAX_SPANS, TC_MAX_READY, TC_PHASES = 8, 4, 8, 32, 16, 8
TC_PHASE_NAMES = ("thresholds", "pilot", "main", "finalize", "exchange")

COMM_ALL_REDUCE = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p)
COMM_ALL_GATHER = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)
COMM_ALL_TO_ALL = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p)


class Comm(ctypes.Structure):
    """``cmh_comm``: the transport of the exchange steps (built-in NCCL, or caller-supplied functions)"""
    _fields_ = [("ctx", ctypes.c_void_p), ("rank", ctypes.c_int32), ("world", ctypes.c_int32),
                ("all_reduce_u32", COMM_ALL_REDUCE), ("all_gather", COMM_ALL_GATHER), ("all_to_all", COMM_ALL_TO_ALL)]


class TcOpts(ctypes.Structure):
    """``cmh_tc_opts``"""
    _fields_ = [("n_pilot", ctypes.c_int32), ("pilot_rows", ctypes.c_int64 * TC_MAX_STAGES), ("prefix", ctypes.c_int32),
                ("n_prefix", ctypes.c_int32), ("prefix_frac", ctypes.c_double * TC_MAX_CUTS),
                ("prefix_min_rows", ctypes.c_int64), ("tighten", ctypes.c_int32), ("cap", ctypes.c_int32),
                ("seg_cap", ctypes.c_int32), ("exact_thresholds", ctypes.c_int32), ("gather", ctypes.c_int32),
                ("n_ready", ctypes.c_int32), ("ready_rows", ctypes.c_int64 * TC_MAX_READY), ("sigma", ctypes.c_double)]


class TcSearch(ctypes.Structure):
    """``cmh_tc_search``"""
    _fields_ = [("bits", ctypes.c_int32), ("K", ctypes.c_int32), ("world", ctypes.c_int32), ("rank", ctypes.c_int32),
                ("nq", ctypes.c_int64), ("nd", ctypes.c_int64), ("nd_total", ctypes.c_int64), ("n_sample", ctypes.c_int64),
                ("n_sample_all", ctypes.c_int64), ("n_stripes", ctypes.c_int32),
                ("stripe_row", ctypes.c_int64 * TC_MAX_STRIPES), ("stripe_index", ctypes.c_int64 * TC_MAX_STRIPES),
                ("n_stages", ctypes.c_int32), ("stage_rows", ctypes.c_int64 * TC_MAX_STAGES),
                ("stage_rows_all", ctypes.c_int64 * TC_MAX_STAGES), ("n_prefix_cuts", ctypes.c_int32),
                ("prefix_cut", ctypes.c_int64 * TC_MAX_CUTS), ("lockstep", ctypes.c_int32), ("n_spans", ctypes.c_int32),
                ("span_lo", ctypes.c_int64 * TC_MAX_SPANS), ("span_hi", ctypes.c_int64 * TC_MAX_SPANS),
                ("span_index", ctypes.c_int64 * TC_MAX_SPANS), ("span_seg_base", ctypes.c_int32 * TC_MAX_SPANS),
                ("span_n_segs", ctypes.c_int32 * TC_MAX_SPANS), ("seg_total", ctypes.c_int32), ("seg_cap", ctypes.c_int32),
                ("n_thr", ctypes.c_int32), ("thr_limit_slot", ctypes.c_int32), ("thr_final_slot", ctypes.c_int32),
                ("per_rank", ctypes.c_int64), ("exch_width", ctypes.c_int32), ("opts", TcOpts), ("sample_plan", Plan),
                ("off_cand", ctypes.c_uint64), ("off_cnt", ctypes.c_uint64), ("off_aux", ctypes.c_uint64),
                ("off_thr", ctypes.c_uint64), ("off_hist", ctypes.c_uint64), ("off_sample_hist", ctypes.c_uint64),
                ("off_part", ctypes.c_uint64), ("off_recv", ctypes.c_uint64), ("off_flags", ctypes.c_uint64),
                ("off_eval", ctypes.c_uint64), ("off_gather", ctypes.c_uint64), ("workspace_bytes", ctypes.c_uint64)]


def sources() -> List[str]:
    return [os.path.join(CSRC_DIR, s) for s in SOURCES]


def _stale() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    built = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC_DIR, h) for h in os.listdir(CSRC_DIR) if h.endswith((".cuh", ".h"))]
    deps.append(os.path.join(INCLUDE_DIR, "cmh_b200.h"))
    return any(os.path.getmtime(d) > built for d in deps if os.path.isfile(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile ``csrc/*.cu`` for sm_100a into ``libcmh_b200.so`` next to this file (nvcc cross-compiles without a
    GPU).  Skips the compile when the library is newer than every source."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", LIB_PATH, *sources()]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


_lib = None
_lock = threading.Lock()


def lib() -> ctypes.CDLL:
    """The loaded library.  Raises RuntimeError when it has not been built - there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                    "(cmh_b200 has no CPU fallback)")
            L = ctypes.CDLL(LIB_PATH)
            _declare(L)
            if L.cmh_abi_version() != ABI_VERSION:
                raise RuntimeError("libcmh_b200.so ABI version mismatch (rebuild: __graft_entry__.build())")
            sizes = (ctypes.c_int32 * 6)()
            L.cmh_struct_sizes(sizes, 6)
            mine = [ctypes.sizeof(c) for c in (CodeSet, Plan, Comm, TcOpts, TcSearch)]
            if list(sizes)[:5] != mine:
                raise RuntimeError(f"ctypes struct mirrors {mine} do not match the library {list(sizes)[:5]}")
            _lib = L
    return _lib


def _declare(L: ctypes.CDLL) -> None:
    vp, i32, i64, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_uint64
    pcs, ppl = ctypes.POINTER(CodeSet), ctypes.POINTER(Plan)
    L.cmh_abi_version.restype = i32
    L.cmh_last_error.restype = ctypes.c_char_p
    L.cmh_device_info.argtypes = [ctypes.POINTER(i32)] * 3 + [ctypes.POINTER(u64)]
    L.cmh_launch_count.restype = ctypes.c_ulonglong
    L.cmh_measure_popc_peak.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double), vp]
    L.cmh_pack_codes.argtypes = [vp, i32, i64, i32, i64, vp, vp, vp, vp]
    L.cmh_unpack_codes.argtypes = [vp, vp, i64, i32, vp, i64, vp]
    L.cmh_pack_scatter.argtypes = [vp, i32, i64, i32, i64, i32, vp, i64, vp, vp, vp, vp]
    L.cmh_hash_head_pack.argtypes = [vp, i32, i64, i32, i64, i32, vp, vp, i32, vp, i64, vp, vp, vp, vp]
    L.cmh_pack_labels.argtypes = [vp, i32, i64, i32, i64, vp, vp, vp]
    L.cmh_synth_codes.argtypes = [u64, i64, i64, i32, vp, vp]
    L.cmh_hamming_dense.argtypes = [pcs, pcs, i32, vp, i64, vp]
    L.cmh_neighbor_dense.argtypes = [vp, i64, vp, i64, i32, vp, i64, vp]
    L.cmh_eval_plan.argtypes = [i64, i64, i32, i32, i32, i32, ppl]
    L.cmh_eval_plan_design.argtypes = [i64, i64, i32, i32, i32, i32, i32, ppl]
    L.cmh_eval_plan_sets.argtypes = [i64, i64, i32, i32, i32, i32, i32, ppl]
    L.cmh_finalize_map_hits.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    L.cmh_eval_hist.argtypes = [ppl, pcs, pcs, vp, vp, vp, vp]
    L.cmh_eval_rank.argtypes = [ppl, pcs, pcs, i64, vp, vp, vp, vp, ctypes.POINTER(i64), i32, vp, vp, vp, vp, vp]
    L.cmh_finalize_map.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    L.cmh_finalize_topn.argtypes = [vp, vp, i64, ctypes.POINTER(i64), i32, i64, vp, vp]
    L.cmh_finalize_pr_workspace_bytes.argtypes = [i64, i32]
    L.cmh_finalize_pr_workspace_bytes.restype = u64
    L.cmh_finalize_pr.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp, vp]
    L.cmh_map_k_workspace_bytes.argtypes = [i64, i64, i32, i32, i32]
    L.cmh_map_k_workspace_bytes.restype = u64
    L.cmh_map_k.argtypes = [pcs, pcs, i32, i32, i64, vp, vp, vp, u64, vp]
    L.cmh_topk.argtypes = [ppl, pcs, pcs, i32, i64, vp, vp, vp]
    L.cmh_topk_merge.argtypes = [vp, i32, i64, i32, vp, vp]
    L.cmh_tc_supported.argtypes = [i32, i32]
    L.cmh_tc_set_workers.argtypes = [i32]
    L.cmh_tc_plan.argtypes = [i64, i64, i32, ctypes.POINTER(i32)]
    L.cmh_tc_collect.argtypes = [vp, i64, vp, i64, i32, i64, vp, i32, i32, i32, i32, vp, vp, vp, vp]
    L.cmh_tc_cand_hist.argtypes = [vp, vp, i64, i32, i32, i32, i32, i32, vp, vp, vp]
    L.cmh_tc_choose.argtypes = [vp, vp, i64, i32, i64, i64, i32, ctypes.c_double, vp, vp, vp]
    L.cmh_tc_choose_prefix.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    L.cmh_tc_choose_seen.argtypes = [vp, vp, i64, i32, i32, vp, vp, vp]
    L.cmh_topk_verify.argtypes = [vp, vp, i64, i32, i64, vp, vp, vp]
    L.cmh_tc_probe.argtypes = [vp, i64, vp, i64, i32, vp, i32, i32, vp, vp, vp, i32, vp]
    L.cmh_topk_threshold.argtypes = [vp, i64, i32, i64, i64, i32, vp, vp]
    L.cmh_topk_finalize.argtypes = [vp, vp, vp, i64, i32, i32, i32, i64, i32, i32, vp, vp, vp, vp]
    L.cmh_topk_merge_verify.argtypes = [vp, i32, i64, i32, i32, i64, vp, vp, vp, vp]
    pcm = ctypes.POINTER(Comm)
    L.cmh_comm_create.argtypes = [i32, ctypes.POINTER(i32), ctypes.POINTER(pcm)]
    L.cmh_comm_unique_id.argtypes = [vp]
    L.cmh_comm_create_rank.argtypes = [vp, i32, i32, ctypes.POINTER(pcm)]
    L.cmh_comm_create_loopback.argtypes = [i32, i32, ctypes.POINTER(pcm)]
    L.cmh_comm_destroy.argtypes = [pcm]
    L.cmh_topk_sharded.argtypes = [pcm, ppl, pcs, pcs, i32, i64, vp, vp, vp, vp]
    L.cmh_map_k_sharded_workspace_bytes.argtypes = [i32, i64, i64, i32, i32, i32, i32]
    L.cmh_map_k_sharded_workspace_bytes.restype = u64
    L.cmh_map_k_sharded.argtypes = [pcm, pcs, pcs, i32, i32, i32, i64, i64, ctypes.POINTER(i64), i32, vp, vp, vp, vp, vp, vp,
                                    vp, u64, vp]
    L.cmh_tc_default_opts.argtypes = [ctypes.POINTER(TcOpts)]
    L.cmh_tc_default_opts.restype = None
    L.cmh_tc_pilot_stages.argtypes = [i64, i64, i32, ctypes.POINTER(i64)]
    L.cmh_struct_sizes.argtypes = [ctypes.POINTER(ctypes.c_int32), i32]
    L.cmh_tc_search_plan.argtypes = [pcm, i64, i64, i64, i32, i32, i32, ctypes.POINTER(i64), ctypes.POINTER(i64), i64,
                                     ctypes.POINTER(TcOpts), ctypes.POINTER(TcSearch)]
    L.cmh_tc_timing_create.argtypes = [ctypes.POINTER(vp)]
    L.cmh_tc_timing_destroy.argtypes = [vp]
    L.cmh_tc_timing_read.argtypes = [vp, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), ctypes.POINTER(i32)]
    L.cmh_tc_timing_launches.argtypes = [vp, ctypes.POINTER(ctypes.c_float), i32]
    L.cmh_topk_tc.argtypes = [ctypes.POINTER(TcSearch), pcm, vp, vp, vp, ctypes.POINTER(vp), vp, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in ("cmh_last_error", "cmh_finalize_pr_workspace_bytes", "cmh_map_k_workspace_bytes",
                        "cmh_launch_count", "cmh_map_k_sharded_workspace_bytes", "cmh_tc_default_opts"):
            fn.restype = i32


def last_error() -> str:
    msg = lib().cmh_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc: int, what: str = "") -> None:
    """0 -> ok; negative -> ValueError (argument / shape errors); positive -> RuntimeError (CUDA error code)."""
    if rc == 0:
        return
    msg = f"{what}: {last_error()} (code {rc})" if what else f"{last_error()} (code {rc})"
    if rc < 0:
        raise ValueError(msg)
    raise RuntimeError(msg)


def i64_array(values) -> Optional[ctypes.Array]:
    values = [int(v) for v in values]
    return (ctypes.c_int64 * len(values))(*values) if values else None
