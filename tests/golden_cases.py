"""Registry of the golden-vector cases shared by `tests/golden/make_golden.py` (generator, needs
/root/reference) and the parity tests (consumers, need only the committed .npz files).

Inputs are never stored: every case regenerates them from its seed with `cmh_b200.synth` (numpy PCG64, stable
across platforms).  Outputs stored per case: the reference's per-query AP for each ``k``, its mAP scalar, and
the first ``topk`` entries of its stable ranking.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cmh_b200.synth import CONFIGS, EvalShape, make_case, make_set_case  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@dataclass(frozen=True)
class GoldenCase:
    name: str
    shape: EvalShape
    ks: Tuple[Optional[int], ...] = (None,)
    topk: int = 64                      # length of the stored ranking prefix
    clustered: bool = True
    ternary_frac: float = 0.0
    zero_query_frac: float = 0.02
    direction: str = "i2t"              # which code pair: i2t -> (q_img, r_txt), t2i -> (q_txt, r_img)
    golden_queries: Optional[int] = None  # only the first n queries are run through the reference
    slow: bool = False                  # config-shaped; minutes on the reference CPU path

    def tensors(self) -> Dict[str, np.ndarray]:
        t = make_case(self.shape, clustered=self.clustered, zero_query_frac=self.zero_query_frac,
                      ternary_frac=self.ternary_frac)
        q, r = ("q_img", "r_txt") if self.direction == "i2t" else ("q_txt", "r_img")
        return {"qB": t[q], "rB": t[r], "qL": t["q_lab"], "rL": t["r_lab"]}

    @property
    def n_golden(self) -> int:
        return self.shape.n_query if self.golden_queries is None else min(self.golden_queries, self.shape.n_query)

    @property
    def path(self) -> str:
        return os.path.join(GOLDEN_DIR, self.name + ".npz")


def _s(name, q, d, bits, nlab, p, seed):
    return EvalShape(name, q, d, bits, nlab, p, None, (), seed)


CASES: Tuple[GoldenCase, ...] = (
    # --- small seeded cases (every code length / label width the reference's trainers produce) ---
    GoldenCase("small_b64_l24", _s("small_b64_l24", 48, 3000, 64, 24, 0.15, 11), ks=(None, 50, 10_000)),
    GoldenCase("small_b16_l21", _s("small_b16_l21", 40, 4000, 16, 21, 0.10, 12), ks=(None, 7), direction="t2i"),
    GoldenCase("small_b32_l21", _s("small_b32_l21", 40, 2500, 32, 21, 0.10, 13), ks=(None, 100)),
    GoldenCase("small_b128_l80", _s("small_b128_l80", 33, 2777, 128, 80, 0.04, 14), ks=(None, 500)),
    GoldenCase("small_b64_uniform", _s("small_b64_uniform", 32, 5000, 64, 24, 0.15, 15), clustered=False),
    GoldenCase("small_b64_ternary", _s("small_b64_ternary", 32, 2000, 64, 24, 0.15, 16), ks=(None, 40),
               ternary_frac=0.05),
    GoldenCase("small_b20_odd", _s("small_b20_odd", 24, 1500, 20, 24, 0.15, 17), ks=(None, 33)),
    GoldenCase("small_b48_odd", _s("small_b48_odd", 24, 1500, 48, 24, 0.15, 18), direction="t2i"),
    GoldenCase("small_b256_long", _s("small_b256_long", 12, 1200, 256, 24, 0.15, 19), ks=(None, 64)),
    GoldenCase("small_b2048_long", _s("small_b2048_long", 6, 700, 2048, 24, 0.15, 20)),
    GoldenCase("small_b64_l291", _s("small_b64_l291", 20, 1800, 64, 291, 0.01, 21), ks=(None, 25)),
    GoldenCase("small_b64_ragged", _s("small_b64_ragged", 131, 1029, 64, 24, 0.15, 22), ks=(None,)),
    GoldenCase("small_b16_ternary", _s("small_b16_ternary", 17, 900, 16, 21, 0.10, 23), ternary_frac=0.15),
    # --- BASELINE.json config shapes (prefix of the queries through the reference; full database) ---
    GoldenCase("c1_full", CONFIGS["c1"], ks=(None,), zero_query_frac=0.01, slow=True),
    GoldenCase("c2_64_prefix", CONFIGS["c2-64"], ks=(None,), zero_query_frac=0.01, golden_queries=96, slow=True),
    GoldenCase("c2_16_prefix", CONFIGS["c2-16"], ks=(None,), zero_query_frac=0.01, golden_queries=64,
               direction="t2i", slow=True),
    GoldenCase("c2_32_prefix", CONFIGS["c2-32"], ks=(None,), zero_query_frac=0.01, golden_queries=64, slow=True),
    GoldenCase("c3_prefix", CONFIGS["c3"], ks=(5000,), zero_query_frac=0.01, golden_queries=96, slow=True),
)

BY_NAME = {c.name: c for c in CASES}
SMALL = tuple(c for c in CASES if not c.slow)


@dataclass(frozen=True)
class SetCase:
    """Set-valued codes (K sub-codes per item) through DPSIH's `mean_average_precision` (train/DPSIH/_utils.py:4-30)."""
    name: str
    n_query: int
    n_db: int
    k_sub: int
    bits: int
    n_labels: int
    label_p: float
    seed: int
    ks: Tuple[Optional[int], ...] = (None,)

    def tensors(self) -> Dict[str, np.ndarray]:
        return make_set_case(self.n_query, self.n_db, self.k_sub, self.bits, self.n_labels, self.label_p, self.seed)

    @property
    def path(self) -> str:
        return os.path.join(GOLDEN_DIR, self.name + ".npz")


SET_CASES: Tuple[SetCase, ...] = (
    SetCase("sets_k2_b64", 40, 3000, 2, 64, 24, 0.15, 41, ks=(None, 100, 7)),
    SetCase("sets_k4_b32", 33, 2111, 4, 32, 21, 0.10, 42, ks=(None, 50)),
    SetCase("sets_k3_b128", 24, 1500, 3, 128, 80, 0.04, 43, ks=(None, 500, 1)),
    SetCase("sets_k1_b16", 30, 4000, 1, 16, 21, 0.10, 44, ks=(None, 64)),
    SetCase("sets_k2_b256", 10, 900, 2, 256, 24, 0.15, 45, ks=(None, 10_000)),
)


def load_golden(case) -> Dict[str, np.ndarray]:
    with np.load(case.path, allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def k_tag(k: Optional[int]) -> str:
    return "all" if k is None else str(int(k))
