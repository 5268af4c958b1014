"""Seeded synthetic hash codes and multi-hot labels at the BASELINE.json shapes.

There is no network, hence no MIRFlickr / NUS-WIDE / MS-COCO data: the retrieval evaluation is exercised on
synthetic codes and labels of exactly the shapes the reference's validation loop produces
(`train/base.py:130-148` -> float32 ``[N, output_dim]`` with entries in {-1, 0, +1};
`dataset/base.py:89-94` -> float32 multi-hot ``[N, nclass]``).

Everything here is numpy PCG64 (`numpy.random.Generator`), whose integer / uniform streams are stable across
platforms and numpy versions, so the GPU box regenerates bit-identical inputs from the seed alone; the golden
fixtures under ``tests/golden`` only have to store *outputs*.

The 100M-row config (C4) cannot be generated as float32 (25.6 GB) - for it `splitmix_rows` defines a
counter-based packed-code generator that the CUDA library implements on the device (`cmh_synth_codes`) and that
can be re-derived slice-wise on the CPU for the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, Optional, Tuple

import numpy as np

__all__ = [
    "EvalShape", "CONFIGS", "make_labels", "make_codes_uniform", "make_codes_clustered", "make_case", "make_set_case",
    "splitmix64", "splitmix_rows",
]


@dataclass(frozen=True)
class EvalShape:
    """One row of SURVEY.md section 8's config table."""
    name: str
    n_query: int
    n_db: int
    bits: int
    n_labels: int
    label_p: float
    k: Optional[int] = None           # calc_map_k's k (None -> mAP@ALL)
    topn: Tuple[int, ...] = ()        # p_topK's K list
    seed: int = 0

    def scaled(self, n_query: Optional[int] = None, n_db: Optional[int] = None, **kw) -> "EvalShape":
        return replace(self, n_query=n_query or self.n_query, n_db=n_db or self.n_db, **kw)

    @property
    def pairs(self) -> int:
        return self.n_query * self.n_db


_TOPN = (1, 100, 200, 300, 400, 500, 600, 700, 800, 900, 1000)

CONFIGS: Dict[str, EvalShape] = {
    # MIRFlickr-25K shape, DCHMT 64-bit, mAP@ALL
    "c1": EvalShape("c1-mirflickr-64b", 2000, 18015, 64, 24, 0.15, None, (), 1000),
    # NUS-WIDE shape, 16/32/64-bit, mAP@ALL + precision@N
    "c2-16": EvalShape("c2-nuswide-16b", 2100, 193734, 16, 21, 0.10, None, _TOPN, 2016),
    "c2-32": EvalShape("c2-nuswide-32b", 2100, 193734, 32, 21, 0.10, None, _TOPN, 2032),
    "c2-64": EvalShape("c2-nuswide-64b", 2100, 193734, 64, 21, 0.10, None, _TOPN, 2064),
    # MS-COCO shape, 128-bit, mAP@5000 + PR curve, 80 labels -> two 64-bit mask words
    "c3": EvalShape("c3-coco-128b", 5000, 117218, 128, 80, 0.04, 5000, (), 3000),
    # large-scale top-1000 (codes come from splitmix_rows, labels unused)
    "c4": EvalShape("c4-large-64b", 1_000_000, 100_000_000, 64, 0, 0.0, 1000, (), 4000),
    # eval stage of the end-to-end valid loop
    "c5": EvalShape("c5-valid-100k-64b", 5000, 100_000, 64, 24, 0.15, None, (), 5000),
}


def make_labels(rng: np.random.Generator, n: int, n_labels: int, p: float, zero_frac: float = 0.0) -> np.ndarray:
    """Bernoulli(p) multi-hot float32 ``[n, n_labels]``; a ``zero_frac`` share of rows is cleared
    (a query without labels is skipped by the reference but still counted in the divisor,
    `utils/calc_utils.py:27-29,38`)."""
    lab = (rng.random((n, n_labels), dtype=np.float32) < np.float32(p)).astype(np.float32)
    if zero_frac > 0.0 and n > 0:
        n_zero = max(1, int(round(n * zero_frac)))
        rows = rng.choice(n, size=min(n, n_zero), replace=False)
        lab[rows] = 0.0
    return lab


def make_codes_uniform(rng: np.random.Generator, n: int, bits: int) -> np.ndarray:
    """i.i.d. fair +-1 codes - the worst case for ties (distances ~ Binomial(bits, 1/2))."""
    return (rng.integers(0, 2, size=(n, bits), dtype=np.int8) * 2 - 1).astype(np.float32)


def make_codes_clustered(rng: np.random.Generator, labels: np.ndarray, prototypes: np.ndarray,
                         flip: float = 0.2) -> np.ndarray:
    """Codes that carry label information: sign(sum of the row's label prototypes), ties and label-free rows
    filled from a private random code, then ``flip`` of the bits inverted.  Items sharing a label end up closer
    than chance, so mAP differs from the base rate and a wrong ranking moves it."""
    n, bits = labels.shape[0], prototypes.shape[1]
    drive = labels @ prototypes                                   # [n, bits] integer-valued float32
    filler = make_codes_uniform(rng, n, bits)
    codes = np.where(drive > 0, 1.0, np.where(drive < 0, -1.0, filler)).astype(np.float32)
    flips = rng.random((n, bits), dtype=np.float32) < np.float32(flip)
    codes[flips] *= -1.0
    return codes


def make_case(shape: EvalShape, *, clustered: bool = True, zero_query_frac: float = 0.01,
              ternary_frac: float = 0.0, seed: Optional[int] = None) -> Dict[str, np.ndarray]:
    """All tensors one `valid()` call feeds the path (`train/base.py:246-262`): image and text codes for the
    query and the retrieval set plus the two label matrices.

    ``ternary_frac`` > 0 overwrites that share of code entries with exact 0.0 - what `torch.sign` emits for a
    zero activation (`train/base.py:141,143`) - to exercise the half-integer distances of SURVEY hard part 3.
    """
    rng = np.random.default_rng(shape.seed if seed is None else seed)
    q_lab = make_labels(rng, shape.n_query, shape.n_labels, shape.label_p, zero_query_frac)
    r_lab = make_labels(rng, shape.n_db, shape.n_labels, shape.label_p, 0.0)
    out = {"q_lab": q_lab, "r_lab": r_lab}
    if clustered:
        proto = make_codes_uniform(rng, shape.n_labels, shape.bits)
    for name, lab in (("q_img", q_lab), ("q_txt", q_lab), ("r_img", r_lab), ("r_txt", r_lab)):
        if clustered:
            codes = make_codes_clustered(rng, lab, proto, 0.2 if name.endswith("img") else 0.25)
        else:
            codes = make_codes_uniform(rng, lab.shape[0], shape.bits)
        if ternary_frac > 0.0:
            codes[rng.random(codes.shape, dtype=np.float32) < np.float32(ternary_frac)] = 0.0
        out[name] = codes
    return out


# ------------------------------------------------------------------------------------------------------------
# counter-based packed codes for the 100M-row config
# ------------------------------------------------------------------------------------------------------------
_M64 = (1 << 64) - 1


def splitmix64(x: np.ndarray) -> np.ndarray:
    """SplitMix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def splitmix_rows(seed: int, row0: int, n: int, words: int, bits: int) -> np.ndarray:
    """Packed codes ``uint64 [n, words]`` for global rows ``row0 .. row0+n``: word ``w`` of row ``r`` is
    ``splitmix64(seed * 0x100000001B3 + r * words + w)``, with the bits at positions >= ``bits`` of the last
    word cleared.  `cmh_synth_codes` in the CUDA library computes the same function."""
    with np.errstate(over="ignore"):
        ctr = (np.arange(row0, row0 + n, dtype=np.uint64)[:, None] * np.uint64(words)
               + np.arange(words, dtype=np.uint64)[None, :])
        base = np.uint64((seed * 0x100000001B3) & _M64)
        out = splitmix64(ctr + base)
    tail = bits - 64 * (words - 1)
    if tail < 64:
        out[:, words - 1] &= np.uint64((1 << tail) - 1)
    return out


# ------------------------------------------------------------------------------------------------------------
# set-valued codes (K sub-codes per item: the input of DPSIH's evaluation, train/DPSIH/_utils.py:4-30)
# ------------------------------------------------------------------------------------------------------------
def make_set_case(n_query: int, n_db: int, k_sub: int, bits: int, n_labels: int, label_p: float, seed: int,
                  zero_query_frac: float = 0.05) -> Dict[str, np.ndarray]:
    """``qB [n_query, k_sub, bits]``, ``rB [n_db, k_sub, bits]`` (+-1 float32) and the two multi-hot label matrices.
    Every sub-code of an item is a differently perturbed copy of the item's clustered code, so the best of the
    ``k_sub x k_sub`` pairs really differs from any fixed pair."""
    rng = np.random.default_rng(seed)
    q_lab = make_labels(rng, n_query, n_labels, label_p, zero_query_frac)
    r_lab = make_labels(rng, n_db, n_labels, label_p, 0.0)
    proto = make_codes_uniform(rng, n_labels, bits)
    out = {"qL": q_lab, "rL": r_lab}
    for name, lab in (("qB", q_lab), ("rB", r_lab)):
        subs = [make_codes_clustered(rng, lab, proto, 0.15 + 0.05 * j) for j in range(k_sub)]
        out[name] = np.ascontiguousarray(np.stack(subs, axis=1))
    return out
