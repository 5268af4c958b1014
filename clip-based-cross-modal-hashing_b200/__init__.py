"""cmh_b200 - B200-native retrieval evaluation (Hamming / mAP@K / precision@N / PR curve / top-K) behind the
call signatures of the reference's `utils/calc_utils.py`.  Import as ``cmh_b200`` (see ../cmh_b200/__init__.py).

    from cmh_b200.calc_utils import calc_map_k_matrix as calc_map_k      # drop-in for train/base.py:11

Sub-modules: ``calc_utils`` (drop-in API), ``utils`` (aliases of the reference's duplicate API), ``dpsih_utils`` (set-valued codes: `train/DPSIH/_utils.py`), ``engine``
(packed sets + kernel wrappers), ``codes`` (binarise-at-source packed code buffers), ``export`` (the reference's `.mat` result files), ``sharded`` (multi-GPU
exchange), ``index`` (resident database for top-K retrieval), ``synth`` (seeded synthetic inputs), ``_cabi`` (ctypes binding of include/cmh_b200.h).
Heavy sub-modules are imported lazily so that `cmh_b200.synth` works without torch / CUDA.
"""
__version__ = "0.1.0"

_API = ("calc_hammingDist", "calc_map_k_matrix", "calc_map_k", "calc_neighbor", "p_topK", "pr_curve", "topk_hamming")


def __getattr__(name):
    if name in _API:
        from . import calc_utils
        return getattr(calc_utils, name)
    raise AttributeError(f"module 'cmh_b200' has no attribute {name!r}")
