"""Signature-compatible aliases of the reference's duplicate evaluation API in ``utils/utils.py`` (never called by
any trainer, but exported by ``utils/__init__.py:1``): same maths as `calc_utils`, except that `calc_map_k`
binarises its inputs itself (`utils/utils.py:77-78`) and `calcHammingDist` accepts numpy (`:105-118`)."""
from __future__ import annotations

import numpy as np
import torch

from . import calc_utils as _cu


def calc_map_k(qB, rB, query_label, retrieval_label, k=None, rank=0):
    """`utils/utils.py:71-102`: ``torch.sign`` is applied to the codes first; the result stays a 0-d float32
    tensor on the device the reference would leave it on (``rank``)."""
    res = _cu.map_k_detail(qB, rB, query_label, retrieval_label, k, rank, binarize=True)
    if res["ap"].shape[0] == 0:
        return 0.0
    return res["map"].reshape(())


def calcHammingDist(B1, B2):
    """`utils/utils.py:105-118`: numpy or torch in, same kind out."""
    as_numpy = isinstance(B1, np.ndarray)
    out = _cu.calc_hammingDist(torch.as_tensor(B1), torch.as_tensor(B2))
    return out.cpu().numpy() if as_numpy else out
