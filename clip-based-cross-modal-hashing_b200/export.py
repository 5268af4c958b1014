"""The on-disk format downstream of the evaluation: the `.mat` dump of `TrainBase.test` / `save_mat`
(`train/base.py:307-322, 328-349`) - one file ``<bits>-ours-<dataset>-<mode>.mat`` under ``<save_dir>/PR_cruve`` with
the float arrays ``q_img, q_txt, r_img, r_txt`` (codes in {-1, 0, +1}) and ``q_l, r_l`` (multi-hot labels) that the
MATLAB PR-curve scripts of the DCMH / DJSRH lineage consume.  Codes may arrive as float tensors (as in the reference)
or packed (`codes.CodeBuffer`, `engine.PackedSet`); packed codes are expanded on the device by `cmh_unpack_codes`.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import check
from .engine import PackedSet, _ptr, _stream

MAT_KEYS = ("q_img", "q_txt", "r_img", "r_txt", "q_l", "r_l")


def unpack_codes(p) -> torch.Tensor:
    """float32 [n, bits] of {-1, 0, +1} on the device from packed codes (inverse of `calc_utils.pack_codes`)."""
    if hasattr(p, "packed") and hasattr(p, "put"):
        p = p.packed()
    if not isinstance(p, PackedSet):
        raise TypeError("unpack_codes takes a PackedSet or a CodeBuffer")
    out = torch.empty((p.n, p.bits), dtype=torch.float32, device=p.device)
    if p.n:
        with torch.cuda.device(p.device):
            check(_cabi.lib().cmh_unpack_codes(_ptr(p.sign), _ptr(p.valid), p.n, p.bits, _ptr(out), p.bits, _stream(p.device)),
                  "cmh_unpack_codes")
    return out


def _codes_numpy(x) -> np.ndarray:
    if isinstance(x, PackedSet) or (hasattr(x, "packed") and hasattr(x, "put")):
        x = unpack_codes(x)
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()                  # train/base.py:307-310
    return np.asarray(x)


def mat_path(save_dir: str, output_dim: int, dataset: str, mode_name: str) -> str:
    """``<save_dir>/<bits>-ours-<dataset>-<mode>.mat`` (`train/base.py:322,348`)."""
    return os.path.join(save_dir, f"{output_dim}-ours-{dataset}-{mode_name}.mat")


def save_mat(query_img, query_txt, retrieval_img, retrieval_txt, query_labels, retrieval_labels, save_dir: str,
             output_dim: Optional[int] = None, dataset: str = "dataset", mode_name: str = "i2t") -> str:
    """Write the reference's result file (`train/base.py:328-349`; ``save_dir`` is the directory that holds it, the
    reference uses ``<args.save_dir>/PR_cruve``).  Returns the path."""
    import scipy.io as scio
    arrays = [_codes_numpy(x) for x in (query_img, query_txt, retrieval_img, retrieval_txt)]
    labels = [x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in (query_labels, retrieval_labels)]
    if output_dim is None:
        output_dim = arrays[0].shape[1]
    os.makedirs(save_dir, exist_ok=True)
    path = mat_path(save_dir, int(output_dim), dataset, mode_name)
    scio.savemat(path, dict(zip(MAT_KEYS, arrays + labels)))
    return path


def load_mat(path: str) -> dict:
    """The six arrays of a result file written by the reference or by `save_mat`."""
    import scipy.io as scio
    m = scio.loadmat(path)
    return {k: m[k] for k in MAT_KEYS}
