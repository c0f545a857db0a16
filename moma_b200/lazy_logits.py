"""Lazy InfoNCE logits handle.

``MoCo.forward`` must return ``logits [B, K+1]`` (MoMA/mem_moco.py:100) that work
with ``nn.CrossEntropyLoss()(logits, labels)`` (helper/loops_moma.py:322,332) and
``accuracy(logits, labels)`` (learning/util.py:25-41, ``.topk(1, 1, True, True)``).
The fused kernel already produced the per-row losses, d loss/d q and the
argmax-is-positive flags in one pass, so the handle answers those two consumers
from the kernel's results and the B x (K+1) matrix is never written to HBM.  Any
other use materialises the dense logits with an explicit kernel (escape hatch).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch
import torch.nn.functional as F

_METADATA = {
    "size", "dim", "numel", "ndimension", "is_contiguous", "element_size", "is_floating_point",
    "type", "__repr__", "__str__", "__format__", "__len__", "stride", "storage_offset", "is_complex",
    "requires_grad_", "__get__", "data_ptr", "__hash__", "is_cuda", "get_device", "nelement",
}


class LazyLogits(torch.Tensor):
    """A [B, K+1] (or [K+1] when B == 1, the reference's .squeeze()) tensor that is not stored."""

    @staticmethod
    def __new__(cls, shape, device, nce, labels, materialize):
        t = torch.Tensor._make_wrapper_subclass(cls, tuple(shape), dtype=torch.float32, device=device,
                                                requires_grad=False)
        t._loss = nce.loss                  # [] mean of the loss rows, differentiable w.r.t. q
        t._rows = nce.rows                  # [B] loss rows, differentiable w.r.t. q
        t._pos_is_max = nce.pos_is_max      # [B] int32
        t._max_logit = nce.max_logit        # [B] max logit per row (for topk values)
        t._acc = nce.acc                    # [1] 100 * mean(pos_is_max)
        t._labels = labels                  # the all-zero labels returned with this handle
        t._materialize = materialize        # callable -> dense [B, K+1] tensor
        t._dense = None
        return t

    # ------------------------------------------------------------------ helpers
    def materialize(self) -> torch.Tensor:
        """Dense logits exactly as mem_moco.py:29-49 lays them out."""
        if self._dense is None:
            self._dense = self._materialize()
        return self._dense

    @property
    def loss_rows(self) -> torch.Tensor:
        return self._rows

    @property
    def pos_is_max(self) -> torch.Tensor:
        return self._pos_is_max

    @property
    def top1_accuracy(self) -> torch.Tensor:
        """[1] tensor, 100 * mean(argmax == 0): learning/util.py:25-41 for the all-zero labels."""
        return self._acc

    def _targets_are_zero(self, target) -> bool:
        # the labels tensor handed out together with this handle is all zeros by construction
        return target is self._labels

    def _cross_entropy(self, target, weight=None, size_average=None, ignore_index=-100, reduce=None,
                       reduction="mean", label_smoothing=0.0):
        if (weight is not None or size_average is not None or reduce is not None or label_smoothing != 0.0
                or not self._targets_are_zero(target) or reduction not in ("mean", "sum", "none")):
            return None
        if reduction == "mean":
            return self._loss
        if reduction == "sum":
            return self._rows.sum()
        return self._rows

    def _topk(self, k, dim=-1, largest=True, sorted=True):
        if k != 1 or not largest or self.dim() != 2 or dim not in (1, -1):
            return None
        # index 0 where the positive wins; otherwise the winning negative's column is not
        # tracked by the fused kernel and reported as -1 (never equal to a valid label)
        idx = torch.where(self._pos_is_max.bool(), 0, -1).to(torch.int64).unsqueeze(1)
        return torch.return_types.topk((self._max_logit.unsqueeze(1), idx))

    # ---------------------------------------------------------------- dispatch
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        lazy = next((a for a in args if isinstance(a, LazyLogits)), None)
        if lazy is not None:
            if func is F.cross_entropy and args and args[0] is lazy:
                out = lazy._cross_entropy(*args[1:], **kwargs)
                if out is not None:
                    return out
            elif func in (torch.topk, torch.Tensor.topk) and args[0] is lazy:
                out = lazy._topk(*args[1:], **kwargs)
                if out is not None:
                    return out
            elif name in _METADATA or name in ("shape", "dtype", "device", "requires_grad", "ndim", "grad",
                                               "grad_fn", "is_leaf", "layout", "names", "_version"):
                with torch._C.DisableTorchFunctionSubclass():
                    return func(*args, **kwargs)

        def dense(a):
            return a.materialize() if isinstance(a, LazyLogits) else a

        args = tuple(dense(a) for a in args)
        kwargs = {k: dense(v) for k, v in kwargs.items()}
        return func(*args, **kwargs)

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        def dense(a):
            return a.materialize() if isinstance(a, LazyLogits) else a
        args = torch.utils._pytree.tree_map(dense, args)
        kwargs = torch.utils._pytree.tree_map(dense, kwargs or {})
        return func(*args, **kwargs)

    def __repr__(self):
        return f"LazyLogits(shape={tuple(self.shape)}, device={self.device})"
