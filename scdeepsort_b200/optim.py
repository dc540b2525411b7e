"""Loss and optimiser of the training loop on the C ABI: ``cross_entropy_sum`` = CrossEntropyLoss(reduction=
'sum') (/root/reference/train.py:36,82) with its gradient from the same launch, and ``Adam`` = torch.optim.Adam
(train.py:34-35,85) as a fused kernel per parameter.  ``Adam`` subclasses ``torch.optim.Optimizer`` and keeps
torch's state layout (``step``, ``exp_avg``, ``exp_avg_sq``), so ``optimizer.state_dict()`` saved next to the
model (train.py:117-123) has the reference's structure."""
import ctypes

import torch

from . import _lib
from .ops import _ptr, _stream


class _CrossEntropySum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels):
        if not (logits.is_cuda and logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1):
            raise ValueError("cross_entropy_sum: expected CUDA fp32 row-major logits")
        labels = labels.to(torch.int64).contiguous()
        m, k = logits.shape
        n_partial = max(1, min((m + 7) // 8, 148 * 8))
        partial = torch.empty(n_partial, device=logits.device, dtype=torch.float32)
        d = torch.empty_like(logits, memory_format=torch.contiguous_format) if logits.requires_grad else None
        lib = _lib.load()
        _lib.check(lib.wsage_softmax_ce(_ptr(logits), logits.stride(0), _ptr(labels), m, k, _ptr(d), d.stride(0) if d is not None else 0,
                                        _ptr(partial), n_partial, _stream()), "wsage_softmax_ce")
        ctx.save_for_backward(d)
        return partial.sum()

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None


def cross_entropy_sum(logits, labels):
    return _CrossEntropySum.apply(logits, labels)


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        lib = _lib.load()
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise ValueError("scdeepsort_b200.optim.Adam: parameters must be contiguous CUDA fp32 tensors")
                st = self.state[p]
                if not st:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] += 1
                g = p.grad.contiguous()
                _lib.check(lib.wsage_adam_step(_ptr(p), _ptr(g), _ptr(st["exp_avg"]), _ptr(st["exp_avg_sq"]), p.numel(),
                                               group["lr"], b1, b2, group["eps"], group["weight_decay"], int(st["step"]),
                                               _stream()), "wsage_adam_step")
