"""The twelve sibling fingerprint kernels: the oracle restatement is pinned against the
unmodified reference (tests/golden/sibling_vectors.npz, oracle/make_golden_siblings.py) on the CPU; the CUDA
path is checked against the same vectors on the GPU.  Bit-exact except Otsuka (sqrt: 2 ulp)."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle.siblings import NAMES, sibling_sim

META = json.load(open(os.path.join(GOLDEN_DIR, "sibling_vectors.json")))
ARR = np.load(os.path.join(GOLDEN_DIR, "sibling_vectors.npz"))
CASES = META["cases"]
# the reference's own KAT expectations ([1,0,1,1] vs [1,1,0,0]; tests/test_kernels/test_fingerprint_kernels/test_*_kernel.py)
KAT = {"dice": 2 / 5, "braun_blanquet": 1 / 3, "faith": 1 / 4, "forbes": 4 / 5, "inner_product": 1.0,
       "intersection": 1.0, "otsuka": 1 / np.sqrt(5), "rand": 1 / 4, "rogers_tanimoto": 1 / 7,
       "russell_rao": 1 / 4, "sogenfrei": 1 / 5, "sokal_sneath": 1 / 7}
EXC = {"RuntimeError": RuntimeError, "ValueError": ValueError}


def _cid(c):
    return f'{c["input"]}-{c["kernel"]}'


def _close(name, got, ref, compute=np.float64):
    if name == "otsuka":   # sqrt: torch's CPU vector sqrt and IEEE sqrt.rn differ in the last ulp of the COMPUTE dtype
        tol = 2.5e-7 if np.dtype(compute) == np.float32 else 5e-16
        return got.dtype == ref.dtype and np.all(np.abs(got - ref) <= tol * np.maximum(np.abs(ref), 1.0) * 2)
    return got.dtype == ref.dtype and np.array_equal(got, ref)


@pytest.mark.parametrize("case", CASES, ids=_cid)
def test_sibling_oracle_matches_reference(case):
    x1, x2 = ARR[f'in_{case["input"]}_x1'], ARR[f'in_{case["input"]}_x2']
    if "raises" in case:
        with pytest.raises((RuntimeError, ValueError)):
            sibling_sim(case["kernel"], x1, x2)
            if case["raises"] == "RuntimeError":       # numpy broadcasts where torch refuses: the product checks shapes
                raise RuntimeError("shape quirk")
        return
    ref = ARR[f'out_{case["input"]}_{case["kernel"]}']
    got = sibling_sim(case["kernel"], x1, x2)
    assert got.shape == ref.shape and _close(case["kernel"], got, ref, x1.dtype)


@pytest.mark.parametrize("name", NAMES)
def test_sibling_kat_values(name):
    a = np.array([[1, 0, 1, 1]], np.float64)
    b = np.array([[1, 1, 0, 0]], np.float64)
    assert np.allclose(sibling_sim(name, a, b), KAT[name], atol=1e-5)


def test_sibling_surface_without_device():
    import gauche_b200.kernels.fingerprint_kernels as fk
    from gauche_b200._kernel_base import ScaleKernel

    for name in NAMES:
        cls_name = "".join(p.capitalize() for p in name.split("_")) + "Kernel"
        cls = getattr(fk, cls_name)
        k = cls()
        assert cls.is_stationary is False and cls.has_lengthscale is False and len(k.state_dict()) == 0
        x = torch.randint(0, 2, (10, 5))
        assert ScaleKernel(cls())(x).size() == torch.Size([10, 10])          # lazy, never evaluated
        mod = __import__(f"gauche_b200.kernels.fingerprint_kernels.{name}_kernel", fromlist=["x"])
        fn = getattr(mod, f"batch_{name}_sim")
        with pytest.raises(ValueError, match="batch dimension"):
            fn(torch.ones(3), torch.ones((2, 3)))


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=_cid)
def test_sibling_cuda_matches_reference(case):
    from gauche_b200.kernels.fingerprint_kernels.sibling_kernels import SIBLINGS

    dev = torch.device("cuda:0")
    fn, cls = SIBLINGS[case["kernel"]]
    x1 = torch.from_numpy(ARR[f'in_{case["input"]}_x1']).to(dev)
    x2 = torch.from_numpy(ARR[f'in_{case["input"]}_x2']).to(dev)
    if "raises" in case:
        with pytest.raises(EXC[case["raises"]]):
            fn(x1, x2)
        return
    ref = ARR[f'out_{case["input"]}_{case["kernel"]}']
    got = fn(x1, x2)
    assert got.is_cuda
    got = got.cpu().numpy()
    assert got.shape == ref.shape and _close(case["kernel"], got, ref, ARR[f'in_{case["input"]}_x1'].dtype), np.abs(got.astype(np.float64) - ref).max()
    via_class = cls().forward(x1, x2).cpu().numpy()
    assert np.array_equal(via_class, got)


@pytest.mark.gpu
def test_sibling_cuda_larger_random_against_oracle():
    from gauche_b200.kernels.fingerprint_kernels.sibling_kernels import SIBLINGS

    dev = torch.device("cuda:0")
    rng = np.random.default_rng(4)
    x = (rng.random((300, 2048)) < 0.03).astype(np.float32)
    y = (rng.random((300, 2048)) < 0.03).astype(np.float32)
    xt, yt = torch.from_numpy(x).to(dev), torch.from_numpy(y).to(dev)
    for name, (fn, _) in SIBLINGS.items():
        got = fn(xt, yt).cpu().numpy()
        assert _close(name, got, sibling_sim(name, x, y), np.float32), name
    from gauche_b200.kernels.fingerprint_kernels import sibling_kernels
    assert sibling_kernels.last_launch["sibling"] == "gk_sibling_gram"      # fused tensor-core epilogue, no second pass


EXT_META = json.load(open(os.path.join(GOLDEN_DIR, "sibling_vectors_ext.json")))
EXT_ARR = np.load(os.path.join(GOLDEN_DIR, "sibling_vectors_ext.npz"))


@pytest.mark.gpu
@pytest.mark.parametrize("case", EXT_META["cases"], ids=_cid)
def test_sibling_cuda_batched_real_and_last_dim_is_batch(case):
    """The reference's surface beyond 2-D integer inputs (tests/golden/sibling_vectors_ext.npz, generated from the
    unmodified reference): leading batch dims incl. broadcasting, real-valued inputs, last_dim_is_batch through the
    kernel classes.  Integer-valued cases bit-exact (Otsuka: sqrt ulp), real-valued to summation-order tolerance; where
    the reference raises, so does the drop-in."""
    from gauche_b200.kernels.fingerprint_kernels import sibling_kernels
    from gauche_b200.kernels.fingerprint_kernels.sibling_kernels import SIBLINGS

    dev = torch.device("cuda:0")
    fn, cls = SIBLINGS[case["kernel"]]
    x1n, x2n = EXT_ARR[f'in_{case["input"]}_x1'], EXT_ARR[f'in_{case["input"]}_x2']
    x1, x2 = torch.from_numpy(x1n).to(dev), torch.from_numpy(x2n).to(dev)
    call = (lambda: cls().covar_dist(x1, x2, last_dim_is_batch=True)) if case["last_dim_is_batch"] else (lambda: fn(x1, x2))
    if "raises" in case:
        with pytest.raises(EXC[case["raises"]]):
            call()
        return
    ref = EXT_ARR[f'out_{case["input"]}_{case["kernel"]}']
    got = call().cpu().numpy()
    assert got.shape == ref.shape and got.dtype == ref.dtype, (got.shape, ref.shape, got.dtype, ref.dtype)
    if case["input"].startswith("real"):
        assert sibling_kernels.last_launch["sibling"] == "gk_sibling_gram_real"
        rtol = 2e-5 if x1n.dtype == np.float32 else 1e-12
        np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol)
    else:
        assert sibling_kernels.last_launch["sibling"] == "gk_sibling_gram"
        assert _close(case["kernel"], got, ref, x1n.dtype), np.abs(got.astype(np.float64) - ref).max()
