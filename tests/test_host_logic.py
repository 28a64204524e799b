"""Host-side logic that needs no GPU: the C-ABI library loads and exports every declared symbol,
the drop-in classes keep the reference's surface, argument errors are raised before any device
work, and the product path fails loudly (no fallback) when CUDA is absent."""
import ctypes
import os
import re

import pytest
import torch

import gauche_b200 as gb
from gauche_b200 import _kernel_base, _lib, ops, screening
from gauche_b200.packed import padded_k


def test_library_exports_every_symbol_of_the_header():
    lib = _lib.load()
    header = open(_lib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(gk_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(raw, name) is not None
    assert isinstance(lib.gk_last_error(), bytes)


def test_query_without_device():
    assert _lib.query(_lib.GK_Q_ABI_VERSION) == 3
    assert _lib.query(_lib.GK_Q_UMMA_TILE_M) == 128 and _lib.query(_lib.GK_Q_UMMA_TILE_N) == 256
    assert _lib.query(_lib.GK_Q_K_ALIGN) == 128
    with pytest.raises(_lib.GaucheB200Error, match="unknown selector"):
        _lib.query(99)


def test_argument_errors_surface_with_messages():
    lib = _lib.load()
    rc = lib.gk_tanimoto_gram_popc(None, 4, None, 3, None, 4, None, 3, 4, None, 3, _lib.GK_F32, 1e-6, None)
    assert rc == -1 and "null pointer" in _lib.last_error()
    rc = lib.gk_tanimoto_gram_umma2(16, 128, 16, 3, 16, 128, 16, 3, 100, 16, 3, _lib.GK_F32, 1e-6, 2, None)
    assert rc == -2 and "multiple of 128" in _lib.last_error()
    rc = lib.gk_tanimoto_gram_bits(16, 3, 16, 3, 16, 4, 16, 3, 3, 16, 3, _lib.GK_F32, 1e-6, None)
    assert rc == -2 and "16-byte aligned" in _lib.last_error()
    rc = lib.gk_tanimoto_gram_umma2(16, 128, 16, 3, 16, 128, 16, 3, 128, 16, 3, _lib.GK_F32, 1e-6, 3, None)
    assert rc == -1 and "cta_group" in _lib.last_error()
    rc = lib.gk_pack_classify(16, _lib.GK_F32, 3, 10, 10, 16, 4, 16, 100, 16, 16, 16, None)
    assert rc == -2
    assert lib.gk_topk_workspace(1000, 0) == -1 and lib.gk_topk_workspace(1000, 20000) == -1
    assert lib.gk_topk_workspace(1000, 100) >= 100 * 8


def test_padded_k():
    assert [padded_k(d) for d in (0, 1, 128, 129, 2048, 2133)] == [128, 128, 128, 256, 2048, 2176]


def test_kernel_classes_keep_the_reference_surface():
    for cls in (gb.TanimotoKernel, gb.MinMaxKernel):
        k = cls()
        assert cls.is_stationary is False and cls.has_lengthscale is False
        assert len(k.state_dict()) == 0 and len(list(k.parameters())) == 0 and len(list(k.buffers())) == 0
        k.load_state_dict({}, strict=True)
    # import paths of the reference namespace
    from gauche_b200.kernels.fingerprint_kernels.minmax_kernel import MinMaxKernel, batch_minmax_sim  # noqa
    from gauche_b200.kernels.fingerprint_kernels.tanimoto_kernel import TanimotoKernel, batch_tanimoto_sim  # noqa
    from gauche_b200.kernels.fingerprint_kernels import MinMaxKernel as M2, TanimotoKernel as T2  # noqa
    import inspect
    assert list(inspect.signature(batch_tanimoto_sim).parameters) == ["x1", "x2", "eps"]
    assert inspect.signature(batch_tanimoto_sim).parameters["eps"].default == 1e-6
    assert list(inspect.signature(TanimotoKernel.forward).parameters)[:4] == ["self", "x1", "x2", "diag"]
    assert list(inspect.signature(TanimotoKernel.covar_dist).parameters)[:4] == ["self", "x1", "x2", "last_dim_is_batch"]


@pytest.mark.parametrize("cls", [gb.TanimotoKernel, gb.MinMaxKernel])
def test_lazy_shapes_like_reference_tests(cls):
    # test_tanimoto_kernel.py:82-96 / test_minmax_kernel.py:41-53: never evaluated, sizes only
    ScaleKernel = _kernel_base.ScaleKernel
    x = torch.randint(0, 2, (10, 5))
    covar = ScaleKernel(cls())(x)
    assert covar.size() == torch.Size([10, 10])
    batch_x = torch.randint(0, 2, (2, 10, 5))
    covar = ScaleKernel(cls())(batch_x)
    assert covar.size() == torch.Size([2, 10, 10])


def test_diag_mode_needs_no_device():
    x = torch.randint(0, 2, (3, 7, 5)).float()
    for cls in (gb.TanimotoKernel, gb.MinMaxKernel):
        d = cls().forward(x, x, diag=True)
        assert d.shape == (3, 7) and d.dtype == x.dtype and torch.all(d == 1)
        with pytest.raises(AssertionError):
            cls().forward(x, x + 1, diag=True)
        with pytest.raises(AssertionError):
            cls().forward(x, x[:, :5], diag=True)


def test_input_validation_precedes_device_work():
    v, m = torch.ones(4), torch.ones((2, 4))
    with pytest.raises(ValueError, match="Tensors must have a batch dimension"):
        gb.batch_tanimoto_sim(v, m)
    with pytest.raises(ValueError, match="Tensors must have a batch dimension"):
        gb.batch_minmax_sim(m, v)
    with pytest.raises(RuntimeError, match="same dtype"):
        gb.batch_tanimoto_sim(m, m.double())
    with pytest.raises(RuntimeError, match="cdist only supports floating-point"):
        gb.batch_minmax_sim(m.long(), m.long())
    with pytest.raises(NotImplementedError):
        gb.batch_minmax_sim(m.clone().requires_grad_(), m)        # CPU tensors: gradients need CUDA inputs
    with pytest.raises(NotImplementedError):
        gb.batch_tanimoto_sim(m.clone().requires_grad_(), m)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_silent_cpu_fallback():
    m = torch.ones((2, 4))
    with pytest.raises(_lib.GaucheB200Error, match="no CPU or PyTorch fallback"):
        gb.batch_tanimoto_sim(m, m)
    with pytest.raises(_lib.GaucheB200Error, match="no CPU or PyTorch fallback"):
        gb.batch_minmax_sim(m, m)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libgauche_b200.so")
    with pytest.raises(_lib.GaucheB200Error, match="not built"):
        _lib.load()


def test_engine_resolution():
    r = ops.resolve_engine
    assert r("tanimoto", "binary", "auto") == "bxh"
    assert r("tanimoto", "count", "bxh") == "umma2" and r("tanimoto", "binary", "mxf4x1") == "mxf4x1"
    assert r("tanimoto", "count", "auto") == "umma2"
    assert r("tanimoto", "count", "bx") == "umma2"
    assert r("tanimoto", "count", "mxf4x1") == "umma2"
    assert r("tanimoto", "count", "umma2x1") == "umma2x1"
    assert r("tanimoto", "binary", "popc") == "popc"
    assert r("tanimoto", "count", "popc") == "u8"
    assert r("tanimoto", "real", "umma2") == "real"
    assert r("minmax", "count", "umma2") == "u8"
    with pytest.raises(ValueError):
        r("tanimoto", "binary", "triton")


def test_collapse_expanded_batches():
    x = torch.arange(12.0).reshape(3, 4)
    e = x.unsqueeze(0).expand(5, -1, -1)
    c = ops._collapse_expanded(e)
    assert c.shape == (1, 3, 4) and torch.equal(c[0], x)
    r = x.unsqueeze(0).repeat(5, 1, 1)
    assert ops._collapse_expanded(r).shape == (5, 3, 4)


def test_shard_bounds_partition_rows():
    for n, w in ((1_000_000, 8), (10, 3), (5, 8), (0, 4)):
        spans = [screening.shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_merge_topk_is_a_stable_global_selection():
    vals = torch.tensor([[5.0, 3.0, 3.0], [5.0, 4.0, float("-inf")]])
    idx = torch.tensor([[2, 0, 1], [7, 9, -1]])
    v, i = screening.merge_topk(vals, idx, 4)
    assert v.tolist() == [5.0, 5.0, 4.0, 3.0] and i.tolist() == [2, 7, 9, 0]


# ------------------------------------------------------------------ host-side packing (no GPU involved)
@pytest.mark.parametrize("dtype", ["float32", "float64", "uint8"])
@pytest.mark.parametrize("rows,dim", [(1, 1), (5, 37), (64, 2048), (33, 2133), (17, 300), (700, 2048)])
def test_host_pack_bits_matches_numpy_packbits(dtype, rows, dim):
    """gk_host_pack_bits writes the packed-bit wire format (np.packbits little-endian), zero padded rows and exact
    popcounts, and flags blocks that are not exactly 0/1."""
    import ctypes
    import numpy as np
    from gauche_b200 import _lib

    lib = _lib.load()
    code = {"float32": _lib.GK_F32, "float64": _lib.GK_F64, "uint8": _lib.GK_U8}[dtype]
    rng = np.random.default_rng(rows * 7 + dim)
    x = (rng.random((rows, dim + 3)) < 0.2).astype(dtype)[:, :dim]          # strided rows (ld = dim + 3)
    words = ((dim + 127) // 128 * 128) // 32
    bits = np.full((rows, words), 0xFFFFFFFF, dtype=np.uint32)
    pop = np.zeros(rows, dtype=np.int32)
    flag = ctypes.c_int32(7)
    rc = lib.gk_host_pack_bits(x.ctypes.data, code, rows, dim, x.strides[0] // x.itemsize, bits.ctypes.data, words,
                               pop.ctypes.data, ctypes.byref(flag), 3)
    ref = np.packbits(np.ascontiguousarray(x).astype(np.uint8), axis=1, bitorder="little")
    assert rc == 0 and flag.value == 0
    assert np.array_equal(bits.view(np.uint8)[:, :ref.shape[1]], ref)
    assert not bits.view(np.uint8)[:, ref.shape[1]:].any()
    assert np.array_equal(pop, x.sum(1).astype(np.int32))
    for bad in ([2] + ([0.5, float("nan")] if dtype != "uint8" else [])):
        y = np.ascontiguousarray(x).copy()
        y[rows // 2, dim // 2] = bad
        rc = lib.gk_host_pack_bits(y.ctypes.data, code, rows, dim, dim, bits.ctypes.data, words, pop.ctypes.data,
                                   ctypes.byref(flag), 0)
        assert rc == 0 and flag.value == 1, bad
    assert lib.gk_host_pack_bits(x.ctypes.data, 99, rows, dim, dim, bits.ctypes.data, words, pop.ctypes.data,
                                 ctypes.byref(flag), 1) < 0
