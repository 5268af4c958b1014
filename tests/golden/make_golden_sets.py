"""Generate `tests/golden/sets_*.npz` by running the REFERENCE'S OWN `mean_average_precision`
(`/root/reference/train/DPSIH/_utils.py:4-30`, imported by path, `torch.argsort` forced stable) on the seeded
set-valued inputs of `tests/golden_cases.py::SET_CASES`.  Build container only (needs /root/reference):

    python tests/golden/make_golden_sets.py

Stored per case: ``ap_<topk>`` (per-query AP from single-query calls), ``map_<topk>`` (the scalar of one call on
the whole query block) and ``map_unstable_<topk>`` (argsort as shipped - informative only).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from golden_cases import SET_CASES, k_tag  # noqa: E402
from oracle import reference_loader as ref  # noqa: E402


def main() -> None:
    print("torch", torch.__version__, "reference", ref.REFERENCE_ROOT)
    for case in SET_CASES:
        T = {k: torch.from_numpy(v) for k, v in case.tensors().items()}
        out = {}
        for topk in case.ks:
            out[f"ap_{k_tag(topk)}"] = ref.reference_set_ap_per_query(T["qB"], T["rB"], T["qL"], T["rL"], topk).numpy()
            out[f"map_{k_tag(topk)}"] = np.float32(float(ref.reference_set_map(T["qB"], T["rB"], T["qL"], T["rL"], topk)))
            out[f"map_unstable_{k_tag(topk)}"] = np.float32(
                float(ref.reference_set_map(T["qB"], T["rB"], T["qL"], T["rL"], topk, stable=False)))
        np.savez_compressed(case.path, **out)
        print(f"{case.name:20s} q={T['qB'].shape[0]:4d} d={T['rB'].shape[0]:6d} K={T['qB'].shape[1]} bits={T['qB'].shape[2]:4d} "
              f"{ {k: float(v) for k, v in out.items() if k.startswith('map_')} }")


if __name__ == "__main__":
    main()
