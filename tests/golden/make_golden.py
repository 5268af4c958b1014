"""Generate `tests/golden/*.npz` by running the REFERENCE'S OWN code (`/root/reference/utils/calc_utils.py`,
imported by path, `torch.sort` forced stable) on the seeded inputs of `tests/golden_cases.py`.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [--only NAME ...] [--skip-slow]

Stored per case (float32 / int32, a few KB each):
  ap_<k>      per-query AP from single-query `calc_map_k_matrix` calls            (calc_utils.py:16-39)
  map_<k>     the scalar the reference returns for the golden query block          (same, one call)
  map_unstable_<k>  the same call with the sort left as shipped (informative only - not a parity target)
  topk_idx    first `topk` database indices of the stable ranking                  (calc_utils.py:30-31)
  topk_dist   their distances from `calc_hammingDist`                              (calc_utils.py:8-13)
  n_rel       relevant-row count per query                                          (calc_utils.py:26-27)
  dense_*     for the tiny cases, the full calc_hammingDist / calc_neighbor blocks
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from golden_cases import CASES, k_tag  # noqa: E402
from oracle import reference_loader as ref  # noqa: E402


def generate(case) -> None:
    t0 = time.time()
    mod = ref.load()
    T = {k: torch.from_numpy(v) for k, v in case.tensors().items()}
    n = case.n_golden
    qB, qL, rB, rL = T["qB"][:n], T["qL"][:n], T["rB"], T["rL"]
    out = {}
    for k in case.ks:
        out[f"ap_{k_tag(k)}"] = ref.reference_ap_per_query(qB, rB, qL, rL, k).numpy().astype(np.float32)
        out[f"map_{k_tag(k)}"] = np.float32(float(ref.reference_map_k(qB, rB, qL, rL, k, stable=True)))
        out[f"map_unstable_{k_tag(k)}"] = np.float32(float(ref.reference_map_k(qB, rB, qL, rL, k, stable=False)))
    dist = mod.calc_hammingDist(qB, rB)                                  # [n, D] float32
    srt = torch.sort(dist, dim=1, stable=True)
    kk = min(case.topk, rB.shape[0])
    out["topk_idx"] = srt.indices[:, :kk].numpy().astype(np.int32)
    out["topk_dist"] = srt.values[:, :kk].numpy().astype(np.float32)
    out["n_rel"] = (mod.calc_neighbor(qL, rL) > 0).sum(1).numpy().astype(np.int64)
    if not case.slow:
        m = min(n, 8)
        out["dense_dist"] = dist[:m, :256].numpy().astype(np.float32)
        out["dense_neighbor"] = mod.calc_neighbor(qL[:m], rL[:256]).numpy().astype(np.float32)
    np.savez_compressed(case.path, **out)
    print(f"{case.name:22s} q={n:5d} d={rB.shape[0]:7d} bits={rB.shape[1]:4d}  "
          f"{os.path.getsize(case.path) / 1024:7.1f} KB  {time.time() - t0:6.1f}s", flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--skip-slow", action="store_true")
    args = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    print("torch", torch.__version__, "numpy", np.__version__, "reference", ref.REFERENCE_ROOT)
    for case in CASES:
        if args.only and case.name not in args.only:
            continue
        if args.skip_slow and case.slow:
            continue
        generate(case)


if __name__ == "__main__":
    main()
