"""Torch-only exact GP around the drop-in kernels — the CALLER side of the path, for the configs whose callers cannot run
here (GPyTorch / BoTorch are not installed in this image).

It drives the kernels through the protocol GPyTorch uses (``gauche/gp.py:128-226``, SURVEY.md §3.1–3.3):
``ScaleKernel(TanimotoKernel())(x1, x2)`` → lazy tensor → ``to_dense()`` → ``ScaleKernel.forward`` →
``base_kernel.forward`` (the hot path), with GPyTorch's parameterisation and objective:

* ``outputscale = softplus(raw_outputscale)``, ``noise = 1e-4 + softplus(raw_noise)`` (``GaussianLikelihood``'s default
  ``GreaterThan(1e-4)``), constant mean;
* training loss = ``-ExactMarginalLogLikelihood`` = ``-(1/N) log N(y | c, s K + noise I)`` by Cholesky, minimised with
  L-BFGS (``benchmarks/run_benchmark.py:122-143`` uses ``fit_gpytorch_model``: scipy L-BFGS-B over the same parameters).
  The kernel has no hyper-parameters, so every closure re-evaluates the SAME Gram: the operand cache makes that a
  cache hit + one kernel launch;
* prediction = ``gauche/gp.py:197-226``: mean ``c + s K(X*, X) (s K + noise I)^-1 (y - c)``, variance from the ``diag=True``
  path (ones) minus the explained part;
* ``screener()`` hands ``alpha = s (s K + noise I)^-1 (y - c)`` to ``TanimotoScreener`` — the BO screening loop of
  ``notebooks/Bayesian Optimisation Over Molecules.ipynb`` cell 5 as one fused pass per candidate block.

``SGPRMin`` does the same for the sparse-GP caller of config 5 (``InducingPointKernel``: ``base(Z, Z)``, ``base(X, Z)``,
``base(X, X, diag=True)``) with the collapsed Titsias bound.

Everything except the kernel matrices is plain torch (Cholesky, triangular solves): the GP solve is out of scope of this
repository; this module exists so that the path can be exercised and tested the way its callers use it.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

from ._kernel_base import ScaleKernel
from .kernels.fingerprint_kernels import MinMaxKernel, TanimotoKernel

_NOISE_FLOOR = 1e-4


def _dense(k) -> torch.Tensor:
    """``kernel(x1, x2)`` returns a lazy tensor (GPyTorch's or the stand-in's): evaluate it."""
    for name in ("to_dense", "evaluate"):
        if hasattr(k, name):
            return getattr(k, name)()
    return k


class ExactGPMin(torch.nn.Module):
    """Exact GP regression with ``ScaleKernel(base_kernel)``, constant mean and homoskedastic Gaussian noise."""

    def __init__(self, train_x: torch.Tensor, train_y: torch.Tensor, base_kernel: Optional[torch.nn.Module] = None):
        super().__init__()
        if train_x.ndim != 2 or train_y.ndim != 1 or train_x.shape[0] != train_y.shape[0]:
            raise ValueError("ExactGPMin expects train_x [N, D] and train_y [N]")
        self.train_x = train_x
        self.train_y = train_y
        self.covar_module = ScaleKernel(base_kernel if base_kernel is not None else TanimotoKernel())
        self.covar_module.to(train_x.device)
        dt = train_y.dtype
        self.raw_noise = torch.nn.Parameter(torch.zeros((), dtype=dt, device=train_y.device))
        self.mean_constant = torch.nn.Parameter(torch.zeros((), dtype=dt, device=train_y.device))
        self.kernel_evaluations = 0

    # ------------------------------------------------------------------ parameters (GPyTorch's transforms)
    @property
    def outputscale(self) -> torch.Tensor:
        return self.covar_module.outputscale.to(self.train_y.dtype)

    @property
    def noise(self) -> torch.Tensor:
        return _NOISE_FLOOR + torch.nn.functional.softplus(self.raw_noise)

    # ------------------------------------------------------------------ the hot path, as the caller reaches it
    def _k(self, x1: torch.Tensor, x2: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.kernel_evaluations += 1
        return _dense(self.covar_module(x1, x2)).to(self.train_y.dtype)

    def _train_chol(self) -> Tuple[torch.Tensor, torch.Tensor]:
        n = self.train_x.shape[0]
        sigma = self._k(self.train_x) + self.noise * torch.eye(n, dtype=self.train_y.dtype, device=self.train_y.device)
        chol = torch.linalg.cholesky(sigma)
        resid = (self.train_y - self.mean_constant).unsqueeze(-1)
        return chol, resid

    def neg_mll(self) -> torch.Tensor:
        """``-ExactMarginalLogLikelihood(likelihood, model)(model(train_x), train_y)`` (divided by N like GPyTorch's)."""
        chol, resid = self._train_chol()
        n = self.train_x.shape[0]
        sol = torch.cholesky_solve(resid, chol)
        quad = (resid * sol).sum()
        logdet = 2.0 * torch.log(torch.diagonal(chol)).sum()
        return 0.5 * (quad + logdet + n * math.log(2.0 * math.pi)) / n

    def fit(self, max_iter: int = 50) -> float:
        """L-BFGS on (raw_outputscale, raw_noise, mean); returns the final loss.  One Gram evaluation per closure."""
        params = [p for p in self.parameters() if p.requires_grad]
        opt = torch.optim.LBFGS(params, max_iter=max_iter, line_search_fn="strong_wolfe", tolerance_grad=1e-9,
                                tolerance_change=1e-12)
        last = [float("nan")]

        def closure():
            opt.zero_grad()
            loss = self.neg_mll()
            loss.backward()
            last[0] = float(loss.detach())
            return loss

        opt.step(closure)
        return last[0]

    # ------------------------------------------------------------------ posterior (gauche/gp.py:197-226)
    @torch.no_grad()
    def alpha(self) -> torch.Tensor:
        """``s (s K + noise I)^-1 (y - c)``: predictive mean = ``c + K_base(X*, X) @ alpha``."""
        chol, resid = self._train_chol()
        return (self.outputscale * torch.cholesky_solve(resid, chol)).squeeze(-1)

    @torch.no_grad()
    def predict(self, x: torch.Tensor, observation_noise: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """Posterior mean and variance at ``x [B, D]`` (one 2-D block: the cross-kernel ``K(X*, X)`` of SURVEY.md §3.2)."""
        chol, resid = self._train_chol()
        k_star = self._k(x, self.train_x)                                   # [B, N] = s * K_base(X*, X)
        mean = self.mean_constant + (k_star @ torch.cholesky_solve(resid, chol)).squeeze(-1)
        prior_var = self.covar_module(x, diag=True).to(self.train_y.dtype)   # forward(diag=True): s * ones
        v = torch.linalg.solve_triangular(chol, k_star.t(), upper=False)    # [N, B]
        var = (prior_var - (v * v).sum(0)).clamp_min(0)
        if observation_noise:
            var = var + self.noise
        return mean, var

    @torch.no_grad()
    def screener(self, **kwargs):
        """``TanimotoScreener`` for this model's posterior mean (Tanimoto base kernel only): candidates stream through the
        fused kernel, ``K(X*, X) @ alpha + c`` never materialises the Gram."""
        from .screening import TanimotoScreener

        if not isinstance(self.covar_module.base_kernel, TanimotoKernel):
            raise TypeError("the fused screening kernel implements the Tanimoto similarity")
        return TanimotoScreener(self.train_x, self.alpha().float(), bias=float(self.mean_constant.detach()), **kwargs)


class SGPRMin(torch.nn.Module):
    """Sparse GP regression with FROZEN inducing fingerprints (config 5): the collapsed Titsias bound and
    its predictive equations on the three kernel calls ``InducingPointKernel`` makes — ``base(Z, Z)``, ``base(X, Z)`` and
    ``base(X, X, diag=True)``.  ``inducing`` are rows of real fingerprints and stay fixed (as learnable parameters they would
    leave the 0/1 domain and take the real-valued kernel)."""

    def __init__(self, train_x: torch.Tensor, train_y: torch.Tensor, inducing: torch.Tensor,
                 base_kernel: Optional[torch.nn.Module] = None, jitter: float = 1e-6):
        super().__init__()
        if train_x.ndim != 2 or inducing.ndim != 2 or train_y.ndim != 1 or train_x.shape[0] != train_y.shape[0]:
            raise ValueError("SGPRMin expects train_x [N, D], train_y [N], inducing [M, D]")
        self.train_x, self.train_y, self.inducing = train_x, train_y, inducing
        self.covar_module = ScaleKernel(base_kernel if base_kernel is not None else TanimotoKernel())
        self.covar_module.to(train_x.device)
        dt = train_y.dtype
        self.raw_noise = torch.nn.Parameter(torch.zeros((), dtype=dt, device=train_y.device))
        self.mean_constant = torch.nn.Parameter(torch.zeros((), dtype=dt, device=train_y.device))
        self.jitter = float(jitter)

    @property
    def outputscale(self) -> torch.Tensor:
        return self.covar_module.outputscale.to(self.train_y.dtype)

    @property
    def noise(self) -> torch.Tensor:
        return _NOISE_FLOOR + torch.nn.functional.softplus(self.raw_noise)

    def _factors(self):
        dt, dv = self.train_y.dtype, self.train_y.device
        m = self.inducing.shape[0]
        kzz = _dense(self.covar_module(self.inducing)).to(dt) + self.jitter * torch.eye(m, dtype=dt, device=dv)   # base(Z, Z)
        kxz = _dense(self.covar_module(self.train_x, self.inducing)).to(dt)                                      # base(X, Z)
        kdiag = self.covar_module(self.train_x, diag=True).to(dt)                                                # base(X, X, diag)
        sig = torch.sqrt(self.noise)
        lz = torch.linalg.cholesky(kzz)
        a = torch.linalg.solve_triangular(lz, kxz.t(), upper=False) / sig                                        # [M, N]
        lb = torch.linalg.cholesky(a @ a.t() + torch.eye(m, dtype=dt, device=dv))
        err = (self.train_y - self.mean_constant).unsqueeze(-1)
        c = torch.linalg.solve_triangular(lb, a @ err, upper=False) / sig
        return lz, lb, a, c, err, kdiag

    def neg_mll(self) -> torch.Tensor:
        """Negative collapsed bound (Titsias 2009) divided by N."""
        lz, lb, a, c, err, kdiag = self._factors()
        n = self.train_x.shape[0]
        noise = self.noise
        bound = (-0.5 * n * math.log(2.0 * math.pi) - torch.log(torch.diagonal(lb)).sum() - 0.5 * n * torch.log(noise)
                 - 0.5 * (err * err).sum() / noise + 0.5 * (c * c).sum() - 0.5 * kdiag.sum() / noise + 0.5 * (a * a).sum())
        return -bound / n

    @torch.no_grad()
    def predict(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        lz, lb, a, c, err, _ = self._factors()
        dt = self.train_y.dtype
        kzs = _dense(self.covar_module(self.inducing, x)).to(dt)                                                 # base(Z, X*)
        t1 = torch.linalg.solve_triangular(lz, kzs, upper=False)
        t2 = torch.linalg.solve_triangular(lb, t1, upper=False)
        mean = self.mean_constant + (t2.t() @ c).squeeze(-1)
        var = (self.covar_module(x, diag=True).to(dt) - (t1 * t1).sum(0) + (t2 * t2).sum(0)).clamp_min(0)
        return mean, var


__all__ = ["ExactGPMin", "SGPRMin", "MinMaxKernel", "TanimotoKernel"]
