"""Import the reference's own `utils/calc_utils.py` by path - TEST INFRASTRUCTURE.

Works only where `/root/reference` is mounted (the build container); the GPU box does not have it, so nothing
on a `-m gpu` test, `smoke()` or `bench.py` path may call this.  Used by `tests/golden/make_golden.py` to
generate the committed golden vectors and by `tests/test_oracle_golden.py::test_live_reference_*` (skipped when
the mount is absent).

The reference ranks with `torch.sort(hamm)` (`utils/calc_utils.py:31`), whose default is not stable; the parity
contract is the stable ranking (ties -> ascending database index).  `stable_sort_patch` swaps in a stable
`torch.sort` for the duration of a call - the reference source itself is left untouched and is never copied.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os

import torch

REFERENCE_ROOT = os.environ.get("CMH_REFERENCE_ROOT", "/root/reference")
_CALC_UTILS = os.path.join(REFERENCE_ROOT, "utils", "calc_utils.py")


def available() -> bool:
    return os.path.isfile(_CALC_UTILS)


def load():
    """The reference module object (`calc_hammingDist`, `calc_map_k_matrix`, `calc_neighbor`)."""
    if not available():
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    spec = importlib.util.spec_from_file_location("_cmh_reference_calc_utils", _CALC_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def stable_sort_patch():
    """Inside the block `torch.sort(x)` behaves as `torch.sort(x, stable=True)`."""
    original = torch.sort

    def _stable(inp, *args, **kwargs):
        kwargs["stable"] = True
        if args:                               # torch.sort(input, dim, descending)
            kwargs.setdefault("dim", args[0])
            if len(args) > 1:
                kwargs.setdefault("descending", args[1])
        return original(inp, **kwargs)

    torch.sort = _stable
    try:
        yield
    finally:
        torch.sort = original


def reference_map_k(qB, rB, qL, rL, k=None, stable=True):
    mod = load()
    ctx = stable_sort_patch() if stable else contextlib.nullcontext()
    with ctx:
        return mod.calc_map_k_matrix(qB, rB, qL, rL, k)


def reference_ap_per_query(qB, rB, qL, rL, k=None, stable=True) -> torch.Tensor:
    """Per-query AP from the reference itself: one single-query call per row (mAP over one query == its AP)."""
    mod = load()
    out = torch.zeros(qB.shape[0], dtype=torch.float32)
    ctx = stable_sort_patch() if stable else contextlib.nullcontext()
    with ctx:
        for i in range(qB.shape[0]):
            out[i] = float(mod.calc_map_k_matrix(qB[i:i + 1], rB, qL[i:i + 1], rL, k))
    return out


# ---- DPSIH's set-based evaluation (`train/DPSIH/_utils.py`) ------------------------------
_DPSIH_UTILS = os.path.join(REFERENCE_ROOT, "train", "DPSIH", "_utils.py")


def load_dpsih():
    """The reference module holding `mean_average_precision` (set-valued codes)."""
    if not os.path.isfile(_DPSIH_UTILS):
        raise FileNotFoundError(f"reference not mounted at {REFERENCE_ROOT}")
    spec = importlib.util.spec_from_file_location("_cmh_reference_dpsih_utils", _DPSIH_UTILS)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@contextlib.contextmanager
def stable_argsort_patch():
    """Inside the block `torch.argsort(x)` behaves as `torch.argsort(x, stable=True)` (`_utils.py:22` ranks with it)."""
    original = torch.argsort

    def _stable(inp, *args, **kwargs):
        kwargs["stable"] = True
        return original(inp, *args, **kwargs)

    torch.argsort = _stable
    try:
        yield
    finally:
        torch.argsort = original


def reference_set_map(qB, rB, qL, rL, topk=None, stable=True):
    mod = load_dpsih()
    ctx = stable_argsort_patch() if stable else contextlib.nullcontext()
    with ctx:
        return mod.mean_average_precision(qB, rB, qL, rL, topk)


def reference_set_ap_per_query(qB, rB, qL, rL, topk=None) -> torch.Tensor:
    """Per-query AP from the reference itself: one single-query call per row."""
    out = torch.zeros(qB.shape[0], dtype=torch.float32)
    for i in range(qB.shape[0]):
        out[i] = float(reference_set_map(qB[i:i + 1], rB, qL[i:i + 1], rL, topk))
    return out
