"""The caller side of the path (gauche/gp.py:128-226 protocol) on the torch-only exact-GP harness:
marginal likelihood, L-BFGS fit, posterior mean / variance and the screening hand-over, against a CPU fp64 GP built on the
oracle's port of the reference kernels."""
import math

import numpy as np
import pytest
import torch

from oracle.torch_port import minmax_sim_torch, tanimoto_sim_torch

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _ref_gp(kfn, x, y, xs, s, noise, c):
    """CPU fp64 reference: -mll / N, posterior mean and variance for given hyper-parameters."""
    n = x.shape[0]
    sigma = s * kfn(x, x) + noise * torch.eye(n, dtype=torch.float64)
    chol = torch.linalg.cholesky(sigma)
    r = (y - c).unsqueeze(-1)
    sol = torch.cholesky_solve(r, chol)
    nmll = 0.5 * ((r * sol).sum() + 2 * torch.log(torch.diagonal(chol)).sum() + n * math.log(2 * math.pi)) / n
    ks = s * kfn(xs, x)
    mean = c + (ks @ sol).squeeze(-1)
    v = torch.linalg.solve_triangular(chol, ks.t(), upper=False)
    var = s - (v * v).sum(0)
    return float(nmll), mean, var


@pytest.mark.parametrize("kernel", ["tanimoto", "minmax"])
def test_exact_gp_harness_matches_cpu_reference(kernel):
    from gauche_b200 import MinMaxKernel, TanimotoKernel, pack_cache
    from gauche_b200.gp_min import ExactGPMin
    from gauche_b200.synthetic import ecfp_bits, ecfp_counts

    g = torch.Generator().manual_seed(31)
    if kernel == "tanimoto":
        x, xs = ecfp_bits(313, 2048, seed=41).double(), ecfp_bits(79, 2048, seed=42).double()
        base, kfn = TanimotoKernel(), tanimoto_sim_torch
    else:
        x, xs = ecfp_counts(313, 2048, extra=85, seed=43).double(), ecfp_counts(79, 2048, extra=85, seed=44).double()
        base, kfn = MinMaxKernel(), minmax_sim_torch
    w = torch.randn(x.shape[1], generator=g, dtype=torch.float64) * (torch.rand(x.shape[1], generator=g) < 0.1)
    y = x @ w + 0.1 * torch.randn(x.shape[0], generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()

    gp = ExactGPMin(x.to(dev()), y.to(dev()), base_kernel=base)
    s0, n0, c0 = float(gp.outputscale.detach()), float(gp.noise.detach()), float(gp.mean_constant.detach())
    assert abs(s0 - math.log(2.0)) < 1e-6 and abs(n0 - (1e-4 + math.log(2.0))) < 1e-12
    ref0, _, _ = _ref_gp(kfn, x, y, xs, s0, n0, c0)
    loss0 = float(gp.neg_mll().detach())
    assert abs(loss0 - ref0) <= 1e-9 * abs(ref0)

    hits0, evals0 = pack_cache.hits, gp.kernel_evaluations
    loss = gp.fit(max_iter=40)
    assert loss < loss0 - 1e-3                      # the optimiser moved outputscale / noise / mean downhill
    # the parameter-free Gram is re-evaluated on the SAME train tensor in every closure: operand-cache hits
    assert pack_cache.hits - hits0 >= gp.kernel_evaluations - evals0 - 1

    s1, n1, c1 = float(gp.outputscale.detach()), float(gp.noise.detach()), float(gp.mean_constant.detach())
    ref1, ref_mean, ref_var = _ref_gp(kfn, x, y, xs, s1, n1, c1)
    assert abs(loss - ref1) <= 1e-7 * max(1.0, abs(ref1))
    mean, var = gp.predict(xs.to(dev()))
    assert torch.allclose(mean.cpu(), ref_mean, rtol=1e-8, atol=1e-9)
    assert torch.allclose(var.cpu(), ref_var.clamp_min(0), rtol=1e-6, atol=1e-9)
    _, var_obs = gp.predict(xs.to(dev()), observation_noise=True)
    assert torch.allclose(var_obs, var + n1)

    if kernel == "tanimoto":
        # BO screening hand-over: the fused K(X*, X) @ alpha + c pass reproduces the posterior mean (fp32 kernel arithmetic)
        scr = gp.screener(block_rows=64)
        scores = scr.scores(xs.float().to(dev()))
        assert torch.allclose(scores.double().cpu(), ref_mean, rtol=1e-4, atol=1e-4)
        vals, idx = scr.screen(xs.float().to(dev()), k=5)
        top = torch.topk(ref_mean, 5)
        assert idx.cpu().tolist() == top.indices.tolist()
    else:
        with pytest.raises(TypeError):
            gp.screener()


def test_sgpr_harness_matches_cpu_reference():
    """Config 5's caller pattern: K_zz, K_xz and the diag call through the drop-in kernel inside the
    collapsed sparse-GP bound and its predictive equations, against the same equations on the oracle's CPU fp64 kernels."""
    from gauche_b200.gp_min import SGPRMin
    from gauche_b200.synthetic import ecfp_bits

    g = torch.Generator().manual_seed(37)
    lib = ecfp_bits(1500, 2048, seed=51).double()
    z, x, xs = lib[:128], lib[128:], ecfp_bits(60, 2048, seed=52).double()       # frozen inducing rows = real fingerprints
    w = torch.randn(2048, generator=g, dtype=torch.float64) * (torch.rand(2048, generator=g) < 0.1)
    y = x @ w + 0.1 * torch.randn(x.shape[0], generator=g, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    gp = SGPRMin(x.to(dev()), y.to(dev()), z.to(dev()))
    s, noise, c = float(gp.outputscale.detach()), float(gp.noise.detach()), 0.0

    kzz = s * tanimoto_sim_torch(z, z) + 1e-6 * torch.eye(128, dtype=torch.float64)
    kxz = s * tanimoto_sim_torch(x, z)
    n = x.shape[0]
    lz = torch.linalg.cholesky(kzz)
    a = torch.linalg.solve_triangular(lz, kxz.t(), upper=False) / math.sqrt(noise)
    lb = torch.linalg.cholesky(a @ a.t() + torch.eye(128, dtype=torch.float64))
    err = (y - c).unsqueeze(-1)
    cc = torch.linalg.solve_triangular(lb, a @ err, upper=False) / math.sqrt(noise)
    bound = (-0.5 * n * math.log(2 * math.pi) - torch.log(torch.diagonal(lb)).sum() - 0.5 * n * math.log(noise)
             - 0.5 * (err * err).sum() / noise + 0.5 * (cc * cc).sum() - 0.5 * n * s / noise + 0.5 * (a * a).sum())
    ref = float(-bound / n)
    got = float(gp.neg_mll().detach())
    assert abs(got - ref) <= 1e-9 * abs(ref)
    kzs = s * tanimoto_sim_torch(z, xs)
    t1 = torch.linalg.solve_triangular(lz, kzs, upper=False)
    t2 = torch.linalg.solve_triangular(lb, t1, upper=False)
    mean, var = gp.predict(xs.to(dev()))
    assert torch.allclose(mean.cpu(), (t2.t() @ cc).squeeze(-1), rtol=1e-8, atol=1e-9)
    assert torch.allclose(var.cpu(), (s - (t1 * t1).sum(0) + (t2 * t2).sum(0)).clamp_min(0), rtol=1e-7, atol=1e-9)
    # the bound is differentiable in the GP's own parameters (the kernel matrices carry no gradient)
    gp.neg_mll().backward()
    assert gp.raw_noise.grad is not None and torch.isfinite(gp.raw_noise.grad)
