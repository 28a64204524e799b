"""Tanimoto kernel on the B200 path.

Drop-in for gauche/kernels/fingerprint_kernels/tanimoto_kernel.py: same names, signatures,
class attributes, shapes, dtypes and error behaviour; the arithmetic runs in hand-written
sm_100a kernels behind the C ABI (``include/gauche_b200.h``).
"""
import torch

from ..._kernel_base import Kernel
from ...ops import tanimoto_similarity


def batch_tanimoto_sim(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``(<x,y> + eps) / (eps + ||x||^2 + ||y||^2 - <x,y>)`` over the last two dims, clamped at 0.

    x1 ``[..., n, d]``, x2 ``[..., m, d]`` -> ``[..., n, m]``; raises ``ValueError`` when either input
    has fewer than two dims (reference tanimoto_kernel.py:11-46).
    """
    return tanimoto_similarity(x1, x2, eps=eps)


class TanimotoKernel(Kernel):
    """GPyTorch kernel ``k(x, x') = <x,x'> / (|x|^2 + |x'|^2 - <x,x'>)`` (reference :49-124).

    No hyper-parameters, no registered buffers: ``state_dict()`` stays empty so checkpoints of
    reference models load unchanged.
    """

    is_stationary = False
    has_lengthscale = False

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def forward(self, x1, x2, diag=False, **params):
        if diag:
            # reference :86-90 — k(x,x)=1 by definition, even for all-zero rows
            assert x1.size() == x2.size() and (x1 is x2 or torch.equal(x1, x2))
            return torch.ones(*x1.shape[:-2], x1.shape[-2], dtype=x1.dtype, device=x1.device)
        return self.covar_dist(x1, x2, **params)

    def covar_dist(self, x1, x2, last_dim_is_batch=False, **params):
        """All-pairs similarity; with ``last_dim_is_batch`` every feature is its own batch
        (``[..., n, d] -> [..., d, n, m]``, reference :120-124)."""
        if last_dim_is_batch:
            x1 = x1.transpose(-1, -2).unsqueeze(-1)
            x2 = x2.transpose(-1, -2).unsqueeze(-1)
        return batch_tanimoto_sim(x1, x2)
