"""Binarise at the source: the packed replacement of the float code buffers `get_code` fills.

The reference's validation loop (`train/base.py:130-148`) allocates two float32 ``[N, bits]`` buffers per loader,
applies `torch.sign` to every batch of encoder outputs (`:141,143`) and stores the rows at the loader's dataset
`index` (`:145-146`; the loaders shuffle even for evaluation).  The DCHMT variant takes the argmax over ``[n, bits, 2]``
logits and maps class 0 to -1 (`train/base.py:150-158`).  `CodeBuffer` does the same per batch with one kernel
(`cmh_pack_scatter`: sign / argmax + bit-pack + scatter) straight into the packed planes the evaluation kernels read,
so the float buffers never exist (C5: 25.6 MB -> 0.8 MB per modality) and `calc_map_k_matrix` has nothing left to pack:

    img, txt = CodeBuffer(N, bits, dev), CodeBuffer(N, bits, dev)
    for image, text, label, index in loader:                      # train/base.py:135
        img.put(index, model.encode_image(image))                 # :140-141,145
        txt.put(index, model.encode_text(text))                   # :142-143,146
    mAP = calc_map_k_matrix(img, txt, query_labels, retrieval_labels)      # CodeBuffers are accepted as they are
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from . import _cabi
from ._cabi import check
from .engine import _TORCH_DTYPE, PackedSet, _ptr, _stream


class CodeBuffer:
    """Packed hash codes of ``length`` items, filled batch by batch in any row order."""

    def __init__(self, length: int, bits: int, device=None):
        if bits < 1 or bits > _cabi.CMH_MAX_BITS:
            raise ValueError(f"code length {bits} outside [1, {_cabi.CMH_MAX_BITS}]")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(
            "cuda", device) if isinstance(device, int) else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("CodeBuffer lives on a CUDA device (cmh_b200 has no CPU path)")
        self.n, self.bits = int(length), int(bits)
        words = (bits + 63) // 64
        # unwritten rows read as all -1 (sign 0, valid 1), like an argmax head that never fired
        self.sign = torch.zeros((self.n, words), dtype=torch.int64, device=self.device)
        self.valid = torch.zeros((self.n, words), dtype=torch.int64, device=self.device)
        self.valid.fill_(-1)
        tail = bits - 64 * (words - 1)
        if tail < 64:
            self.valid[:, words - 1] = (1 << tail) - 1
        self._counters = torch.zeros(2, dtype=torch.int64, device=self.device)   # (#zeros, #rows with a bad index)

    def _index(self, index, n: int) -> Optional[torch.Tensor]:
        if index is None:
            return None
        idx = torch.as_tensor(np.asarray(index) if not isinstance(index, torch.Tensor) else index)
        idx = idx.to(device=self.device, dtype=torch.int64).contiguous().view(-1)
        if idx.numel() != n:
            raise ValueError(f"{idx.numel()} indices for {n} rows")
        return idx

    def _put(self, index, x: torch.Tensor, mode: int) -> None:
        if x.device != self.device:                  # host tensors and tensors of ANOTHER GPU alike: the kernel runs here
            x = x.to(self.device)
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
            x = x.float()
        x = x.detach()
        want = (self.bits,) if mode == 0 else (self.bits, 2)
        if x.dim() != len(want) + 1 or tuple(x.shape[1:]) != want:
            raise ValueError(f"expected [n, {', '.join(map(str, want))}], got {tuple(x.shape)}")
        x = x.contiguous()
        n = x.shape[0]
        idx = self._index(index, n)
        ld = self.bits * (2 if mode else 1)
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_pack_scatter(_ptr(x), _TORCH_DTYPE[x.dtype], n, self.bits, ld, mode, _ptr(idx), self.n,
                                               _ptr(self.sign), _ptr(self.valid), _ptr(self._counters),
                                               _stream(self.device)), "cmh_pack_scatter")

    def put(self, index, activations: torch.Tensor) -> None:
        """``sign(activations)`` ``[n, bits]`` -> rows ``index`` (`train/base.py:141-146`); None = rows 0..n-1."""
        self._put(index, activations, 0)

    def put_argmax(self, index, logits: torch.Tensor) -> None:
        """DCHMT head: ``argmax(logits [n, bits, 2], -1)`` with class 0 -> -1 (`train/base.py:150-158`)."""
        self._put(index, logits, 1)

    def put_head(self, index, hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                 relu: bool = False) -> None:
        """The whole DCHMT head on the way in (`cmh_hash_head_pack`): ``hidden`` [n, h] are the activations after
        `HashLayer.fc` (`relu=True` applies the reference's `torch.relu`, model/DCHMT.py:22), ``weight`` [bits, 2, h] /
        ``bias`` [bits, 2] the `bits` stacked `nn.Linear(h, 2)` layers; bit j = argmax over the two logits (a tie is
        class 0 = -1, train/base.py:150-158).  The [n, bits, 2] logits and the float codes never exist."""
        x = hidden.detach()
        if x.device != self.device:
            x = x.to(self.device)
        if x.dtype not in (torch.float32, torch.float16, torch.bfloat16):
            x = x.float()
        if x.dim() != 2:
            raise ValueError(f"expected hidden activations [n, h], got {tuple(x.shape)}")
        x = x.contiguous()
        n, h = x.shape
        w = weight.detach().to(device=self.device, dtype=torch.float32).contiguous()
        if tuple(w.shape) != (self.bits, 2, h):
            raise ValueError(f"weight must be [{self.bits}, 2, {h}], got {tuple(w.shape)}")
        b = None
        if bias is not None:
            b = bias.detach().to(device=self.device, dtype=torch.float32).contiguous()
            if tuple(b.shape) != (self.bits, 2):
                raise ValueError(f"bias must be [{self.bits}, 2], got {tuple(b.shape)}")
        idx = self._index(index, n)
        with torch.cuda.device(self.device):
            check(_cabi.lib().cmh_hash_head_pack(_ptr(x), _TORCH_DTYPE[x.dtype], n, h, h, 1 if relu else 0, _ptr(w), _ptr(b),
                                                 self.bits, _ptr(idx), self.n, _ptr(self.sign), _ptr(self.valid),
                                                 _ptr(self._counters), _stream(self.device)), "cmh_hash_head_pack")

    def reset(self) -> None:
        """Forget everything written so far (all rows read as -1 again, zero / bad-index counters cleared): a buffer
        reused across epochs returns to the +-1 fast path even if an earlier epoch stored exact zeros."""
        self.sign.zero_()
        self.valid.fill_(-1)
        tail = self.bits - 64 * (self.sign.shape[1] - 1)
        if tail < 64:
            self.valid[:, self.sign.shape[1] - 1] = (1 << tail) - 1
        self._counters.zero_()

    def recount(self) -> int:
        """Recount the exact zeros actually stored (rows overwritten since may have removed them) and return their
        number; `packed()` then drops the valid plane again when there is none left."""
        words = self.sign.shape[1]
        full = torch.full((words,), -1, dtype=torch.int64, device=self.device)
        tail = self.bits - 64 * (words - 1)
        if tail < 64:
            full[words - 1] = (1 << tail) - 1
        missing = (~self.valid) & full                # real bit positions whose entry is an exact zero
        n_zero = int(sum(int(((missing >> s) & 1).sum()) for s in range(64))) if bool(missing.any()) else 0
        self._counters[0] = n_zero
        return n_zero

    def packed(self) -> PackedSet:
        """The planes as the evaluation kernels take them (one tiny D2H: the zero / bad-index counters)."""
        n_zero, n_bad = (int(v) for v in self._counters.tolist())
        if n_bad:
            raise IndexError(f"{n_bad} rows were put at indices outside [0, {self.n})")
        return PackedSet(self.sign, self.valid if n_zero else None, None, self.n, self.bits, 0, n_zero)
