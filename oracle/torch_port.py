"""torch-CPU port of the reference path (test infrastructure / CPU baseline only).

Same ATen operators, in the same order, as
  gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:33-46
  gauche/kernels/fingerprint_kernels/minmax_kernel.py:29-44
so that timing it on the GPU box's host cores measures what the reference itself would cost
there (the reference is pure Python on torch and cannot travel to the box).  Used for inputs the
exact-integer oracle does not cover (real-valued fingerprints) and as ``cpu_baseline`` in bench.py.
"""
from __future__ import annotations

import torch


def tanimoto_sim_torch(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    cross = x1 @ x2.transpose(-1, -2)
    sq1 = (x1 ** 2).sum(dim=-1, keepdim=True)
    sq2 = (x2 ** 2).sum(dim=-1, keepdim=True)
    ratio = (cross + eps) / (eps + sq1 + sq2.transpose(-1, -2) - cross)
    return ratio.clamp_min_(0)


def minmax_sim_torch(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    l1_a = x1.sum(dim=-1, keepdim=True)
    l1_b = x2.sum(dim=-1, keepdim=True)
    total = l1_a + l1_b.transpose(-1, -2)
    dist = torch.cdist(x1, x2, p=1)
    ratio = (total - dist + eps) / (total + dist + eps)
    return ratio.clamp_min_(0)
