"""The oracle against the golden vectors generated from the reference's own code (tests/golden/make_golden.py),
and the three restatements (sorted torch / counting numpy / counting C) against one another.  CPU only."""
import numpy as np
import pytest
import torch

from golden_cases import BY_NAME, CASES, SMALL, k_tag, load_golden
from oracle import c_oracle, cmh_oracle as orc, reference_loader as ref

TOL = 1e-6


def _keys_from_golden(g):
    d2 = np.rint(g["topk_dist"].astype(np.float64) * 2).astype(np.uint64)
    return (d2 << np.uint64(32)) | g["topk_idx"].astype(np.uint64)


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c.name)
def test_sorted_oracle_matches_reference_goldens(case):
    g, T = load_golden(case), case.tensors()
    n = case.n_golden
    for k in case.ks:
        ap, n_rel = orc.ap_per_query_sorted(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k)
        assert np.array_equal(n_rel.numpy(), g["n_rel"])
        np.testing.assert_allclose(ap.numpy(), g[f"ap_{k_tag(k)}"], rtol=0, atol=TOL)
        m = float(orc.map_k_sorted(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k))
        assert abs(m - float(g[f"map_{k_tag(k)}"])) < TOL
    dist, idx = orc.topk_sorted(T["qB"][:n], T["rB"], case.topk)
    assert np.array_equal(idx.numpy().astype(np.int32), g["topk_idx"])          # bit-exact ranking
    assert np.array_equal(dist.numpy(), g["topk_dist"])
    if "dense_dist" in g:
        m = g["dense_dist"].shape[0]
        assert np.array_equal(orc.hamming_dist(T["qB"][:m], T["rB"][:256]).numpy(), g["dense_dist"])
        assert np.array_equal(orc.neighbor(T["qL"][:m], T["rL"][:256]).numpy(), g["dense_neighbor"])


@pytest.mark.parametrize("case", SMALL, ids=lambda c: c.name)
def test_counting_oracle_matches_goldens(case):
    g, T = load_golden(case), case.tensors()
    n = min(case.n_golden, 16)
    for k in case.ks:
        ap, n_rel = orc.ap_per_query_counting(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k)
        assert np.array_equal(n_rel, g["n_rel"][:n])
        np.testing.assert_allclose(ap, g[f"ap_{k_tag(k)}"][:n], rtol=0, atol=TOL)
    keys = orc.topk_counting(T["qB"][:n], T["rB"], case.topk)
    assert np.array_equal(keys, _keys_from_golden(g)[:n])


@pytest.mark.parametrize("case", CASES, ids=lambda c: c.name)
def test_c_oracle_matches_goldens(case):
    """The C restatement is the only oracle fast enough for the full-size parity runs: pin it on every case,
    including the BASELINE.json config shapes."""
    g, T = load_golden(case), case.tensors()
    n = case.n_golden
    for k in case.ks:
        m, ap, n_rel, _ = c_oracle.map_k(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k)
        assert np.array_equal(n_rel, g["n_rel"])
        np.testing.assert_allclose(ap, g[f"ap_{k_tag(k)}"], rtol=0, atol=TOL)
        assert abs(m - float(g[f"map_{k_tag(k)}"])) < TOL
    qs, qv, _, _ = orc.pack_codes(T["qB"][:n])
    ds, dv, _, _ = orc.pack_codes(T["rB"])
    keys = c_oracle.topk_packed(qs, qv, ds, dv, case.shape.bits, case.topk)
    assert np.array_equal(keys, _keys_from_golden(g))


@pytest.mark.parametrize("name", ["small_b64_l24", "small_b64_ternary", "small_b128_l80"])
def test_precision_and_pr_restatements_agree(name):
    """p_topK / pr_curve are not in the reference (parity unpinned): the sorted and counting forms of the frozen
    definitions must at least agree with each other."""
    case = BY_NAME[name]
    T = case.tensors()
    n = 24
    topn = (1, 5, 100, 1000, 10**6)
    p_sorted = orc.p_topk_sorted(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], topn).numpy()
    _, _, _, p_c = c_oracle.map_k(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], None, topn)
    np.testing.assert_allclose(p_sorted, p_c, rtol=0, atol=TOL)
    P1, R1 = orc.pr_curve_dense(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"])
    P2, R2 = orc.pr_curve_counting(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"])
    np.testing.assert_allclose(P1.numpy(), P2, rtol=0, atol=TOL)
    np.testing.assert_allclose(R1.numpy(), R2, rtol=0, atol=TOL)


def test_known_answers():
    """Hand-checkable cases."""
    B = 16
    # (i) identical codes: all distances 0 -> ranking = index order; relevant rows at 1, 3 -> AP = (1/2 + 2/4) / 2
    q = np.ones((1, B), np.float32); r = np.ones((5, B), np.float32)
    qL = np.array([[1, 0]], np.float32); rL = np.array([[0, 1], [1, 0], [0, 1], [1, 1], [0, 0]], np.float32)
    assert abs(float(orc.map_k_sorted(q, r, qL, rL)) - 0.5) < 1e-7
    assert abs(c_oracle.map_k(q, r, qL, rL)[0] - 0.5) < 1e-12
    # (ii) one relevant row, 3 rows strictly closer -> AP = 1/4
    r2 = np.ones((5, B), np.float32); r2[4, :3] = -1
    rL2 = np.array([[0, 1], [0, 1], [0, 1], [0, 1], [1, 0]], np.float32)
    r2[:4, :1] = -1   # four rows at distance 1, the relevant one at distance 3
    assert abs(float(orc.map_k_sorted(q, r2, qL, rL2)) - 1 / 5) < 1e-7
    # (iii) label-free query is skipped but stays in the divisor
    q3 = np.ones((2, B), np.float32); qL3 = np.array([[1, 0], [0, 0]], np.float32)
    assert abs(float(orc.map_k_sorted(q3, r, qL3, rL)) - 0.25) < 1e-7
    assert abs(c_oracle.map_k(q3, r, qL3, rL)[0] - 0.25) < 1e-12
    # (iv) k < n_rel: only the first k relevant rows; k > D behaves like None
    assert abs(float(orc.map_k_sorted(q, r, qL, rL, 1)) - 0.5) < 1e-7
    assert abs(float(orc.map_k_sorted(q, r, qL, rL, 100)) - 0.5) < 1e-7
    # (v) exact zeros -> half-integer distances
    qz = np.ones((1, B), np.float32); qz[0, 0] = 0
    assert float(orc.hamming_dist(qz, r)[0, 0]) == 0.5
    qs, qv, nz, _ = orc.pack_codes(qz); ds, dv, _, _ = orc.pack_codes(r)
    assert nz == 1 and int(orc.dist2_packed(qs[0], qv[0], ds, dv, B)[0]) == 1


def test_splitmix_known_answers():
    from cmh_b200.synth import splitmix64, splitmix_rows
    # SplitMix64 reference outputs for the seed 0 stream (state 0 -> first output, etc.)
    assert int(splitmix64(np.array([0], np.uint64))[0]) == 0xE220A8397B1DCDAF
    assert int(splitmix64(np.array([0x9E3779B97F4A7C15], np.uint64))[0]) == 0x6E789E6AA1B965F4
    rows = splitmix_rows(4000, 10, 4, 1, 40)
    assert rows.shape == (4, 1) and (rows >> np.uint64(40)).max() == 0
    assert np.array_equal(rows[2:], splitmix_rows(4000, 12, 2, 1, 40))     # slice-wise reproducible


@pytest.mark.skipif(not ref.available(), reason="/root/reference not mounted (GPU box)")
@pytest.mark.parametrize("name", ["small_b64_l24", "small_b64_ternary", "small_b20_odd"])
def test_live_reference_agrees_with_oracle(name):
    """Where the reference is mounted, run its own module (sort forced stable) against the restatements."""
    case = BY_NAME[name]
    T = {k: torch.from_numpy(v) for k, v in case.tensors().items()}
    n = 12
    for k in case.ks:
        want = float(ref.reference_map_k(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k))
        assert abs(float(orc.map_k_sorted(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k)) - want) < TOL
        assert abs(c_oracle.map_k(T["qB"][:n].numpy(), T["rB"].numpy(), T["qL"][:n].numpy(), T["rL"].numpy(), k)[0] - want) < TOL
    mod = ref.load()
    assert torch.equal(mod.calc_hammingDist(T["qB"][:4], T["rB"][:100]), orc.hamming_dist(T["qB"][:4], T["rB"][:100]))


def test_valid_loop_encoder_matches_reference_clip():
    """f2: `cmh_b200.valid_loop.Clip` (fused attention, same parameter names) loaded with the state dict of the
    reference's own `CLIP` class reproduces its `encode_image` / `encode_text` (golden made by
    tests/golden/make_golden_clip.py from /root/reference/model/base/model.py) in float32 within 2e-5."""
    import os
    import numpy as np
    import torch
    from cmh_b200.valid_loop import Clip, ClipConfig
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clip_tiny.npz"))
    cfg = ClipConfig(*[int(v) for v in z["cfg"]])
    model = Clip(cfg).float().eval()
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    with torch.no_grad():
        img = model.encode_image(torch.from_numpy(z["image"]))
        txt = model.encode_text(torch.from_numpy(z["text"]))
    np.testing.assert_allclose(img.numpy(), z["img_feat"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(txt.numpy(), z["txt_feat"], rtol=0, atol=2e-5)


# ---- f4: DPSIH's set-based evaluation (train/DPSIH/_utils.py:4-30) -------------------------------------------------
from golden_cases import SET_CASES  # noqa: E402


@pytest.mark.parametrize("case", SET_CASES, ids=lambda c: c.name)
def test_set_oracle_matches_reference_goldens(case):
    """Both restatements of `mean_average_precision` against vectors generated by the reference's own function
    (tests/golden/make_golden_sets.py, argsort forced stable)."""
    g, T = load_golden(case), case.tensors()
    for k in case.ks:
        ap, hits = orc.set_ap_per_query_sorted(T["qB"], T["rB"], T["qL"], T["rL"], k)
        np.testing.assert_allclose(ap.numpy(), g[f"ap_{k_tag(k)}"], rtol=0, atol=TOL)
        assert abs(float(orc.set_map_sorted(T["qB"], T["rB"], T["qL"], T["rL"], k)) - float(g[f"map_{k_tag(k)}"])) < TOL
        ap2, hits2 = orc.set_ap_per_query_counting(T["qB"], T["rB"], T["qL"], T["rL"], k)
        np.testing.assert_allclose(ap2, g[f"ap_{k_tag(k)}"], rtol=0, atol=TOL)
        assert np.array_equal(hits.numpy(), hits2)


def test_set_oracle_known_answers():
    """Hand-checkable: K = 2 sub-codes of 8 bits; only the BEST pair counts; textbook AP@topk."""
    q = -np.ones((1, 2, 8), np.float32); q[0, 1] = 1.0                        # sub-codes: all -1, all +1
    r = np.ones((4, 2, 8), np.float32)
    r[0, :, :4] = -1                                                          # both sub-codes half/half: best dist 4
    r[1, 0] = -1                                                              # holds an all -1 sub-code: dist 0
    r[2, 1, :2] = -1                                                          # (+1 x8, six +1): best dist 0 via q's +1 code
    r[3, :, :6] = -1                                                          # two -1 x6: best dist min(2, 6) = 2
    qL = np.array([[1, 0]], np.float32)
    rL = np.array([[1, 0], [0, 1], [1, 0], [1, 1]], np.float32)
    # ranking: row1 (0), row2 (0), row3 (2), row0 (4); relevant: rows 0, 2, 3 at ranks 4, 2, 3
    want_all = (1 / 2 + 2 / 3 + 3 / 4) / 3
    assert abs(float(orc.set_map_sorted(q, r, qL, rL)) - want_all) < 1e-7
    assert abs(float(orc.set_map_sorted(q, r, qL, rL, 3)) - (1 / 2 + 2 / 3) / 2) < 1e-7     # rank 4 is cut off
    assert float(orc.set_map_sorted(q, r, qL, rL, 1)) == 0.0                                  # no relevant row in the top 1
    ap, hits = orc.set_ap_per_query_counting(q, r, qL, rL, 3)
    assert hits.tolist() == [2] and abs(ap[0] - (1 / 2 + 2 / 3) / 2) < 1e-12


@pytest.mark.skipif(not ref.available(), reason="/root/reference not mounted (GPU box)")
def test_live_reference_set_map_agrees_with_oracle():
    case = SET_CASES[0]
    T = {k: torch.from_numpy(v) for k, v in case.tensors().items()}
    n = 10
    for k in (None, 25):
        want = float(ref.reference_set_map(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k))
        assert abs(float(orc.set_map_sorted(T["qB"][:n], T["rB"], T["qL"][:n], T["rL"], k)) - want) < TOL
