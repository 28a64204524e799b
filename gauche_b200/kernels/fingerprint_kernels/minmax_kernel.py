"""MinMax kernel on the B200 path.

Drop-in for gauche/kernels/fingerprint_kernels/minmax_kernel.py: same names, signatures,
class attributes, shapes, dtypes and error behaviour; the arithmetic runs in hand-written
sm_100a kernels behind the C ABI (``include/gauche_b200.h``).
"""
import torch

from ..._kernel_base import Kernel
from ...ops import minmax_similarity


def batch_minmax_sim(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``(|x|_1 + |y|_1 - |x-y|_1 + eps) / (|x|_1 + |y|_1 + |x-y|_1 + eps)`` over the last two dims,
    clamped at 0 (reference minmax_kernel.py:10-44); float dtypes only, like ``torch.cdist``."""
    return minmax_similarity(x1, x2, eps=eps)


class MinMaxKernel(Kernel):
    """GPyTorch kernel ``k(x, x') = sum_i min(x_i, x'_i) / sum_i max(x_i, x'_i)`` (reference :47-120)."""

    is_stationary = False
    has_lengthscale = False

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def forward(self, x1, x2, diag=False, **params):
        if diag:
            assert x1.size() == x2.size() and (x1 is x2 or torch.equal(x1, x2))
            return torch.ones(*x1.shape[:-2], x1.shape[-2], dtype=x1.dtype, device=x1.device)
        return self.covar_dist(x1, x2, **params)

    def covar_dist(self, x1, x2, last_dim_is_batch=False, **params):
        if last_dim_is_batch:
            x1 = x1.transpose(-1, -2).unsqueeze(-1)
            x2 = x2.transpose(-1, -2).unsqueeze(-1)
        return batch_minmax_sim(x1, x2)
