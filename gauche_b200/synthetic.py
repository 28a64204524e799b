"""Seeded synthetic fingerprints with the statistics of SURVEY.md §8(d) (no RDKit / datasets here).

ECFP-like bits: per-row target popcount ``n_i ~ clip(round(N(48, 14^2)), 8, 160)``; column
propensities ``w_c ∝ 1/(1+c)^0.6`` over a fixed random permutation of the columns (a few very
common bits, long tail).  Bits are drawn as independent Bernoulli(min(1, n_i * w_c)) so that
million-row libraries can be generated on the device in chunks; the realised popcount is
Poisson-binomial around ``n_i``.  Count fingerprints multiply the support by ``1 + Poisson(0.3)``
and optionally append ``extra`` fragment-count columns ``Poisson(0.5)`` (fragprints, 2048+85).
"""
from __future__ import annotations

import torch


def _column_weights(dim: int, gen: torch.Generator, device) -> torch.Tensor:
    w = 1.0 / (1.0 + torch.arange(dim, dtype=torch.float64, device=device)) ** 0.6
    perm = torch.randperm(dim, generator=gen, device=device)
    w = w[perm]
    return (w / w.sum()).to(torch.float32)


def ecfp_bits(rows: int, dim: int = 2048, seed: int = 0, device="cpu", dtype=torch.float32,
              chunk: int = 65536, adversarial: bool = False) -> torch.Tensor:
    """``[rows, dim]`` dense 0/1 matrix in ``dtype``."""
    device = torch.device(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    w = _column_weights(dim, gen, device)
    out = torch.empty((rows, dim), dtype=dtype, device=device)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        n = torch.randn((r1 - r0, 1), generator=gen, device=device) * 14.0 + 48.0
        n = n.round().clamp_(8, min(160, dim))
        p = (n * w.unsqueeze(0)).clamp_(max=1.0)
        out[r0:r1] = (torch.rand((r1 - r0, dim), generator=gen, device=device) < p).to(dtype)
    if adversarial and rows >= 6:
        out[0] = 0            # all-zero
        out[1] = 1            # all-one
        out[2] = 0
        out[2, dim // 3] = 1  # single bit
        out[3] = out[4]       # duplicate
    return out


def ecfp_counts(rows: int, dim: int = 2048, extra: int = 0, seed: int = 0, device="cpu",
                dtype=torch.float32, chunk: int = 65536) -> torch.Tensor:
    """``[rows, dim+extra]`` count fingerprints, values in ``[0, 255]``."""
    device = torch.device(device)
    bits = ecfp_bits(rows, dim, seed, device, torch.float32, chunk)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 7919)
    out = torch.empty((rows, dim + extra), dtype=dtype, device=device)
    for r0 in range(0, rows, chunk):
        r1 = min(rows, r0 + chunk)
        lam = torch.full((r1 - r0, dim), 0.3, device=device)
        mult = 1.0 + torch.poisson(lam, generator=gen)
        out[r0:r1, :dim] = (bits[r0:r1] * mult).clamp_(max=255).to(dtype)
        if extra:
            lam2 = torch.full((r1 - r0, extra), 0.5, device=device)
            out[r0:r1, dim:] = torch.poisson(lam2, generator=gen).clamp_(max=255).to(dtype)
    return out
