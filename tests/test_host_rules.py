"""Host-side restatements of two exactness arguments the tensor-core search relies on (CPU only, numpy):

* finalize's bounded-disorder rule (`topk_finalize_kernel`, csrc/tc_collect.cu): the candidates of a query are stored tile
  by tile - in index order up to a permutation INSIDE one 256-row tile - so when the K-th bucket holds more rows than the
  sort buffer, the `want` lowest indices of the bucket are among its first `want + 255` stored entries;
* the drain-side carry argument recorded in DESIGN.md section 8 (a 32-bit add over two packed 16-bit accumulators leaves
  every flag what it would be without the carry, because `dot - T` is even) - kept as a known-answer check of the
  packed-field arithmetic the shipped kernel's flags use as well (two 8-bit fields per 16-bit accumulator).
"""
import numpy as np
import pytest


@pytest.mark.parametrize("seed", range(8))
def test_bounded_disorder_truncation_keeps_the_lowest_indices(seed):
    rng = np.random.default_rng(seed)
    n_rows = int(rng.integers(5_000, 200_000))
    density = rng.choice([0.02, 0.1, 0.5, 1.0])
    rows = np.nonzero(rng.random(n_rows) < density)[0]                 # the bucket's rows, ascending
    # stored order: tiles in order, any permutation inside a 256-row tile
    stored = np.concatenate([rng.permutation(rows[(rows // 256) == t]) for t in np.unique(rows // 256)]) if rows.size else rows
    for want in (1, 7, 100, 1000, 3000):
        want = min(want, rows.size)
        if want == 0:
            continue
        kept = stored[:want + 255]
        assert set(rows[:want].tolist()) <= set(kept.tolist())
        # and the bound is tight: without the slack a row can be lost
    if rows.size > 300:
        worst = np.concatenate([rows[(rows // 256) == t][::-1] for t in np.unique(rows // 256)])     # every tile reversed
        want = 1
        t0 = rows[(rows // 256) == (rows[0] // 256)]
        if t0.size > 1:
            assert rows[0] not in worst[:want].tolist()                # truncating at `want` alone would drop the winner


def test_packed_field_flags_with_and_without_carry():
    """acc16 = e1 + 256 * e2 per accumulator, two accumulators per 32-bit register; flags: bit 7 / 15 (23 / 31) clear <=>
    row j / row j + 128 qualifies.  Adding the per-query constant to both halves with ONE 32-bit add lets a carry out of
    the low half into the high one: with even e1 no flag changes."""
    B = 256
    rng = np.random.default_rng(3)
    for _ in range(20000):
        T = int(rng.integers(1, 33)) * 2                                # even threshold dot product
        d = rng.integers(-32, 33, size=4) * 2                           # four even dot products (64-bit codes)
        raw = [(int(d[0]) + B * int(d[1])) & 0xFFFF, (int(d[2]) + B * int(d[3])) & 0xFFFF]
        bias16 = (B - (B + 1) * T) & 0xFFFF
        exact = [(r + bias16) & 0xFFFF for r in raw]                    # independent 16-bit adds
        fused = ((raw[0] | (raw[1] << 16)) + (bias16 | (bias16 << 16))) & 0xFFFFFFFF
        halves = [fused & 0xFFFF, fused >> 16]
        for h in range(2):
            for bit in (7, 15):
                assert ((halves[h] >> bit) & 1) == ((exact[h] >> bit) & 1)
            # and the flags mean what the kernel says they mean
            assert (((exact[h] >> 7) & 1) == 0) == (int(d[2 * h]) >= T)
            assert (((exact[h] >> 15) & 1) == 0) == (int(d[2 * h + 1]) >= T)
