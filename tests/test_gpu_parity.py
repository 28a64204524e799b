"""GPU parity: the CUDA path (through the C ABI) against the oracle and the reference's golden
vectors.  Bar: bit-exact for integer-valued fingerprints (counts AND floats); 1e-6 relative
(fp32) / 1e-12 (fp64) for real-valued inputs, where summation order is the only difference."""
import numpy as np
import pytest
import torch

import oracle
from oracle import ccounts
from conftest import case_id, golden, golden_cases, is_integer_valued

pytestmark = pytest.mark.gpu

REAL_RTOL = {np.dtype("float32"): 2e-6, np.dtype("float64"): 1e-12}


def dev():
    return torch.device("cuda:0")


def _engines_for(case, x1, x2):
    if "minmax" in case["fn"].lower() or not is_integer_valued(x1, x2):
        return [None]
    binary = all(a is None or bool(np.all((a == 0) | (a == 1))) for a in (x1, x2))
    return (["bxh", "mxf4x1", "bx", "umma2", "umma2x1", "popc", "u8"] if binary else ["umma2", "umma2x1", "u8"])


def _cases_with_engines():
    out = []
    for c in golden_cases():
        x1, x2 = golden().inputs(c)
        for e in _engines_for(c, x1, x2):
            out.append(pytest.param(c, e, id=f"{case_id(c)}-{e or 'auto'}"))
    return out


def _run_case(case, engine):
    import gauche_b200 as gb
    from gauche_b200 import ops

    x1, x2 = golden().inputs(case)
    t1 = torch.from_numpy(x1).to(dev())
    t2 = t1 if x2 is None else torch.from_numpy(x2).to(dev())
    fn, kw = case["fn"], case["kwargs"]
    if fn == "batch_tanimoto_sim":
        return ops.tanimoto_similarity(t1, t2, engine=engine)
    if fn == "batch_minmax_sim":
        return gb.batch_minmax_sim(t1, t2)
    if fn == "TanimotoKernel.covar_dist":
        return gb.TanimotoKernel().covar_dist(t1, t2, **kw)
    if fn == "TanimotoKernel.forward":
        return gb.TanimotoKernel().forward(t1, t2, **kw)
    if fn == "MinMaxKernel.covar_dist":
        return gb.MinMaxKernel().covar_dist(t1, t2, **kw)
    if fn == "MinMaxKernel.forward":
        return gb.MinMaxKernel().forward(t1, t2, **kw)
    raise AssertionError(fn)


@pytest.mark.parametrize("case,engine", _cases_with_engines())
def test_golden_reference_vectors(case, engine):
    x1, x2 = golden().inputs(case)
    ref = golden().output(case)
    got_t = _run_case(case, engine)
    assert got_t.is_cuda
    got = got_t.cpu().numpy()
    assert got.dtype == ref.dtype, (got.dtype, ref.dtype)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if is_integer_valued(x1, x2):
        assert np.array_equal(got, ref), f"max abs diff {np.abs(got.astype(np.float64) - ref).max()}"
    else:
        rtol = REAL_RTOL[ref.dtype]
        np.testing.assert_allclose(got, ref, rtol=rtol * 8, atol=rtol)


# ------------------------------------------------------------------ pack kernel vs C oracle
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.int64, torch.int32, torch.uint8,
                                   torch.float16, torch.bfloat16, torch.int8, torch.int16, torch.bool])
@pytest.mark.parametrize("rows,dim", [(1, 1), (5, 37), (33, 128), (17, 2048), (9, 2133)])
def test_pack_matches_c_oracle(dtype, rows, dim):
    from gauche_b200.packed import pack, padded_k

    rng = np.random.default_rng(rows * 1000 + dim)
    if dtype == torch.bool:
        x = rng.random((rows, dim)) < 0.2
    else:
        x = rng.poisson(0.4, size=(rows, dim))
        x[0, 0] = 200
    xt = torch.from_numpy(np.asarray(x)).to(dtype).to(dev())
    p = pack(xt)
    kp = padded_k(dim)
    bits, q8, rs, rq, flags = ccounts.pack(xt.cpu().to(torch.float64).numpy(), kp)
    assert p.kp == kp
    assert np.array_equal(p.bits.cpu().numpy().view(np.uint32), bits)
    if p.kind == "real":      # e.g. negative int8 values: no u8 operand exists, the dense fp path serves these
        with pytest.raises(ValueError):
            p.q8
    else:                     # bits: expanded lazily from the bit words; counts: written by the packer's second pass
        assert np.array_equal(p.q8.cpu().numpy(), q8)
    assert np.array_equal(p.rowsum.cpu().numpy(), rs)
    assert np.array_equal(p.rowsq.cpu().numpy(), rq)
    assert p.flags == flags


def test_pack_classification_and_strides():
    from gauche_b200.packed import pack

    base = torch.zeros((6, 300), dtype=torch.float32, device=dev())
    base[:, ::7] = 1
    assert pack(base).kind == "binary"
    v = base[:, 3:260]                      # row-strided, unaligned view
    pv = pack(v)
    _, q8, rs, _, fl = ccounts.pack(v.cpu().double().numpy(), pv.kp)
    assert pv.kind == "binary" and np.array_equal(pv.q8.cpu().numpy(), q8)
    assert np.array_equal(pv.rowsum.cpu().numpy(), rs)
    t = base.t()                            # inner stride != 1 -> made contiguous
    assert np.array_equal(pack(t).q8.cpu().numpy()[:, :6], t.cpu().numpy().astype(np.uint8))
    c = base.clone(); c[2, 5] = 3
    assert pack(c).kind == "count"
    for bad in (0.5, -1.0, 256.0, float("nan"), float("inf")):
        r = base.clone(); r[1, 1] = bad
        assert pack(r).kind == "real", bad


# ------------------------------------------------------------------ randomized shapes
SHAPES = [(1, 1, 1), (1, 300, 5), (3, 1, 37), (127, 129, 128), (129, 257, 2048), (64, 64, 2133),
          (200, 513, 2048), (385, 100, 96), (130, 260, 4096)]


@pytest.mark.parametrize("n,m,d", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
@pytest.mark.parametrize("engine", ["bxh", "mxf4x1", "bx", "umma2", "umma2x1", "popc", "u8"])
def test_tanimoto_bits_random_shapes(n, m, d, dtype, engine):
    from gauche_b200 import ops

    rng = np.random.default_rng(n * 7 + m * 13 + d)
    x1 = (rng.random((n, d)) < 0.05).astype(np.float64)
    x2 = (rng.random((m, d)) < 0.05).astype(np.float64)
    if n > 2:
        x1[0] = 0
        x1[1] = 1
    t1 = torch.from_numpy(x1).to(dtype).to(dev())
    t2 = torch.from_numpy(x2).to(dtype).to(dev())
    ref = oracle.tanimoto_sim(t1.cpu().numpy(), t2.cpu().numpy())
    got = ops.tanimoto_similarity(t1, t2, engine=engine).cpu().numpy()
    assert np.array_equal(got, ref)
    # every shape — also row pitches the TMA cannot address (m = 1, 129, 257, 513) — runs the engine it names
    assert ops.last_launch["tanimoto"] == {"bx": "gk_tanimoto_gram_bits", "mxf4x1": "gk_tanimoto_gram_mxf4",
                                           "bxh": "gk_tanimoto_gram_hybrid" if dtype == torch.float32 else "gk_tanimoto_gram_mxf4",
                                           "umma2": "gk_tanimoto_gram_umma2", "umma2x1": "gk_tanimoto_gram_umma2",
                                           "popc": "gk_tanimoto_gram_popc", "u8": "gk_tanimoto_gram_u8"}[engine]
    inter, n1, n2 = oracle.tanimoto_counts(x1, x2)
    cnt = ops.intersection_counts(t1, t2, engine=engine).cpu().numpy()
    assert cnt.dtype == np.int32 and np.array_equal(cnt, inter)
    # union counts from the packer's exact norms
    from gauche_b200.packed import pack
    u = pack(t1).rowsq.cpu().numpy()[:, None].astype(np.int64) + pack(t2).rowsq.cpu().numpy()[None, :] - cnt
    assert np.array_equal(u, n1[:, None] + n2[None, :] - inter)


@pytest.mark.parametrize("n,m,d", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_counts_random_shapes(n, m, d, dtype):
    from gauche_b200 import ops

    rng = np.random.default_rng(n * 3 + m * 11 + d)
    x1 = rng.poisson(0.15, size=(n, d)).astype(np.float64)
    x2 = rng.poisson(0.15, size=(m, d)).astype(np.float64)
    x1[0, 0] = 9
    t1 = torch.from_numpy(x1).to(dtype).to(dev())
    t2 = torch.from_numpy(x2).to(dtype).to(dev())
    ref_t = oracle.tanimoto_sim(t1.cpu().numpy(), t2.cpu().numpy())
    for engine in ("umma2", "umma2x1", "u8"):
        got = ops.tanimoto_similarity(t1, t2, engine=engine).cpu().numpy()
        assert np.array_equal(got, ref_t), engine
    ref_m = oracle.minmax_sim(t1.cpu().numpy(), t2.cpu().numpy())
    d_ref, _, _ = oracle.minmax_l1(x1, x2)
    for engine in ("tc", "u8", None):     # "tc": thermometer planes on the tensor cores, any output pitch
        got = ops.minmax_similarity(t1, t2, engine=engine).cpu().numpy()
        assert np.array_equal(got, ref_m), engine
        assert np.array_equal(ops.l1_distances(t1, t2, engine=engine).cpu().numpy(), d_ref), engine


def test_u8_maximum_values_exact():
    """255-valued counts at D=2133: dot products reach 1.4e8 > 2^24; integer accumulators stay exact."""
    from gauche_b200 import ops

    n, m, d = 70, 90, 2133
    rng = np.random.default_rng(5)
    x1 = rng.integers(0, 256, size=(n, d)).astype(np.float64)
    x2 = rng.integers(0, 256, size=(m, d)).astype(np.float64)
    x1[0] = 255
    x2[0] = 255
    t1, t2 = torch.from_numpy(x1).to(dev()), torch.from_numpy(x2).to(dev())
    inter, _, _ = oracle.tanimoto_counts(x1, x2)
    for engine in ("umma2", "umma2x1", "u8"):
        assert np.array_equal(ops.intersection_counts(t1, t2, engine=engine).cpu().numpy(), inter)
        got = ops.tanimoto_similarity(t1, t2, engine=engine).cpu().numpy()
        assert np.array_equal(got, oracle.tanimoto_sim(x1, x2))          # fp64: exact
    d_ref, _, _ = oracle.minmax_l1(x1, x2)
    assert np.array_equal(ops.l1_distances(t1, t2).cpu().numpy(), d_ref)
    assert np.array_equal(ops.minmax_similarity(t1, t2).cpu().numpy(), oracle.minmax_sim(x1, x2))


@pytest.mark.parametrize("dtype,rtol", [(torch.float32, 2e-5), (torch.float64, 1e-12)])
def test_real_valued_against_torch_port(dtype, rtol):
    from gauche_b200 import ops

    g = torch.Generator().manual_seed(1)
    x1 = torch.rand((150, 333), generator=g, dtype=dtype)
    x2 = torch.rand((70, 333), generator=g, dtype=dtype) - 0.3
    for fn, port in ((ops.tanimoto_similarity, oracle.tanimoto_sim_torch), (ops.minmax_similarity, oracle.minmax_sim_torch)):
        got = fn(x1.to(dev()), x2.to(dev())).cpu()
        ref = port(x1, x2)
        assert got.dtype == dtype
        torch.testing.assert_close(got, ref, rtol=rtol, atol=rtol)


# ------------------------------------------------------------------ boundary behaviour
def test_empty_and_degenerate_inputs():
    import gauche_b200 as gb

    x = torch.zeros((0, 16), device=dev())
    y = torch.ones((3, 16), device=dev())
    assert gb.batch_tanimoto_sim(x, y).shape == (0, 3)
    assert gb.batch_tanimoto_sim(y, x).shape == (3, 0)
    assert gb.batch_minmax_sim(x, y).shape == (0, 3)
    z = torch.zeros((4, 16), device=dev())
    assert torch.equal(gb.batch_tanimoto_sim(z, z).cpu(), torch.ones(4, 4))      # eps/eps
    assert torch.equal(gb.batch_minmax_sim(z, z).cpu(), torch.ones(4, 4))
    e = torch.zeros((2, 0), device=dev())
    assert torch.equal(gb.batch_tanimoto_sim(e, e).cpu(), oracle.tanimoto_sim_torch(e.cpu(), e.cpu()))


def test_views_expanded_batches_and_aliasing():
    import gauche_b200 as gb

    g = torch.Generator().manual_seed(2)
    base = (torch.rand((40, 300), generator=g) < 0.1).float()
    xd = base.to(dev())
    ref_full = oracle.tanimoto_sim(base.numpy(), base.numpy())
    assert np.array_equal(gb.batch_tanimoto_sim(xd, xd).cpu().numpy(), ref_full)          # x1 is x2
    v1, v2 = xd[3:30:2, 10:290], xd[5:, 10:290]                                          # strided views
    ref = oracle.tanimoto_sim(base[3:30:2, 10:290].numpy(), base[5:, 10:290].numpy())
    assert np.array_equal(gb.batch_tanimoto_sim(v1, v2).cpu().numpy(), ref)
    # stride-0 expanded train batch, as gauche/gp.py:176-190 builds it
    test = xd[:6].unsqueeze(0).repeat(3, 1, 1)
    test[1] = xd[6:12]
    train = xd.unsqueeze(0).expand(3, -1, -1)
    got = gb.batch_tanimoto_sim(test, train).cpu().numpy()
    exp = oracle.tanimoto_sim(test.cpu().numpy(), np.broadcast_to(base.numpy(), (3, 40, 300)))
    assert got.shape == (3, 6, 40) and np.array_equal(got, exp)


def test_host_tensors_round_trip():
    """CPU tensors in -> CPU tensor out (host buffers through the device path)."""
    import gauche_b200 as gb

    g = torch.Generator().manual_seed(3)
    x = (torch.rand((50, 2048), generator=g) < 0.03).double()
    out = gb.batch_tanimoto_sim(x, x)
    assert out.device.type == "cpu" and out.dtype == torch.float64
    assert np.array_equal(out.numpy(), oracle.tanimoto_sim(x.numpy(), x.numpy()))


def test_error_behaviour_matches_reference():
    import gauche_b200 as gb

    v = torch.ones(4, device=dev())
    m = torch.ones((2, 4), device=dev())
    with pytest.raises(ValueError, match="batch dimension"):
        gb.batch_tanimoto_sim(v, m)
    with pytest.raises(ValueError, match="batch dimension"):
        gb.batch_minmax_sim(m, v)
    with pytest.raises(RuntimeError):
        gb.batch_minmax_sim(m.long(), m.long())          # cdist rejects integer tensors
    with pytest.raises(RuntimeError):
        gb.batch_tanimoto_sim(m, m.double())             # matmul rejects mixed dtypes
    with pytest.raises(AssertionError):
        gb.TanimotoKernel().forward(m, m + 1, diag=True)
    k = gb.TanimotoKernel()
    assert torch.equal(k.forward(m, m, diag=True).cpu(), torch.ones(2))
    assert len(k.state_dict()) == 0 and len(gb.MinMaxKernel().state_dict()) == 0


def test_pack_cache_reuses_operands():
    import gauche_b200 as gb
    from gauche_b200.packed import pack_cache

    pack_cache.clear()
    x = (torch.rand((64, 512), device=dev()) < 0.1).float()
    h0, m0 = pack_cache.hits, pack_cache.misses
    a = gb.batch_tanimoto_sim(x, x)
    b = gb.batch_tanimoto_sim(x, x)
    assert pack_cache.misses == m0 + 1 and pack_cache.hits >= h0 + 1
    assert torch.equal(a, b)
    x[0, :] = 1                                           # in-place edit bumps _version -> repack
    c = gb.batch_tanimoto_sim(x, x)
    assert pack_cache.misses == m0 + 2
    assert np.array_equal(c.cpu().numpy(), oracle.tanimoto_sim(x.cpu().numpy(), x.cpu().numpy()))


# ------------------------------------------------------------------ screening helpers
def test_rowdot_and_topk():
    from gauche_b200 import screening

    g = torch.Generator(device="cuda").manual_seed(4)
    K = torch.rand((257, 1003), generator=g, device=dev())
    alpha = torch.randn((1003,), generator=g, device=dev())
    s = screening.rowdot(K, alpha, bias=0.25)
    torch.testing.assert_close(s, K @ alpha + 0.25, rtol=1e-5, atol=1e-4)
    for n, k in ((1000, 10), (100000, 1000), (5000, 5000), (70000, 16384)):
        sc = torch.randn((n,), generator=g, device=dev())
        sc[::7] = sc[0]                                    # heavy ties
        vals, idx = screening.topk(sc, k, index_base=1000)
        order = torch.sort(sc, descending=True, stable=True)
        assert torch.equal(vals, order.values[:k])
        assert torch.equal(idx, order.indices[:k] + 1000)


# ------------------------------------------------------------------ epilogue division
def test_fast_division_is_ieee():
    """The tensor-core epilogue's branch-free division must equal div.rn.f32 bit for bit on its
    operand domain: every (intersection, union-ish denominator) of 2048/4096-bit fingerprints
    (exhaustive), and random operands over the count-fingerprint range."""
    from gauche_b200 import _lib
    from gauche_b200.packed import stream_ptr

    lib = _lib.load()
    d = dev()
    eps = torch.tensor(1e-6, dtype=torch.float32)

    def check(a, b):
        a = a.contiguous().float(); b = b.contiguous().float()
        fast = torch.empty_like(a); ieee = torch.empty_like(a)
        rc = lib.gk_selftest_fdiv_f32(a.data_ptr(), b.data_ptr(), fast.data_ptr(), ieee.data_ptr(),
                                      a.numel(), stream_ptr(d))
        _lib.check(rc, "gk_selftest_fdiv_f32")
        torch.cuda.synchronize()
        assert torch.equal(fast.view(torch.int32), ieee.view(torch.int32))
        assert torch.equal(ieee, a / b)

    # exhaustive: num = dot + eps, den = (union-ish integer) (+eps effects), dot <= 4096, den <= 8192
    dot = torch.arange(0, 4097, device=d, dtype=torch.float32)
    den = torch.arange(0, 8193, device=d, dtype=torch.float32)
    num = (dot + eps.to(d))[:, None].expand(-1, den.numel())
    for den_variant in (den + eps.to(d), (eps.to(d) + den), den.clamp_min(1.0)):
        check(num.reshape(-1), den_variant[None, :].expand(dot.numel(), -1).reshape(-1))
    # random count-range operands (dot products up to 2133*255^2, denominators up to 2x that)
    g = torch.Generator(device="cuda").manual_seed(9)
    for hi_a, hi_b in ((1.4e8, 2.8e8), (1e4, 1e6), (255.0, 65025.0), (1.0, 4.3e9)):
        a = torch.rand(1 << 22, generator=g, device=d) * hi_a + 1e-6
        b = torch.rand(1 << 22, generator=g, device=d) * hi_b + 1e-6
        check(a.floor() + 1e-6, b.floor() + 1.0)
        check(a, b)


def test_fast_division_is_ieee_fp64():
    """fp64 twin of the test above: the branch-free sequence (nvcc's own fast path without its range test) equals
    __ddiv_rn bit for bit on the epilogue's operand domain."""
    from gauche_b200 import _lib
    from gauche_b200.packed import stream_ptr

    lib = _lib.load()
    d = dev()

    def check(a, b):
        a = a.contiguous().double(); b = b.contiguous().double()
        fast = torch.empty_like(a); ieee = torch.empty_like(a)
        rc = lib.gk_selftest_fdiv_f64(a.data_ptr(), b.data_ptr(), fast.data_ptr(), ieee.data_ptr(),
                                      a.numel(), stream_ptr(d))
        _lib.check(rc, "gk_selftest_fdiv_f64")
        torch.cuda.synchronize()
        assert torch.equal(fast.view(torch.int64), ieee.view(torch.int64))
        assert torch.equal(ieee, a / b)

    for eps in (1e-6, 1e-12, 1.0, 0.37):
        dot = torch.arange(0, 4097, device=d, dtype=torch.float64)
        den = torch.arange(0, 8193, device=d, dtype=torch.float64)
        num = (dot + eps)[:, None].expand(-1, den.numel())
        for den_variant in (den + eps, (eps + den), den.clamp_min(1.0)):
            check(num.reshape(-1), den_variant[None, :].expand(dot.numel(), -1).reshape(-1))
    g = torch.Generator(device="cuda").manual_seed(10)
    for hi_a, hi_b in ((1.4e8, 2.8e8), (1e4, 1e6), (255.0, 65025.0), (1.0, 4.3e9), (3.3e7, 6.7e7)):
        a = torch.rand(1 << 22, generator=g, device=d, dtype=torch.float64) * hi_a + 1e-6
        b = torch.rand(1 << 22, generator=g, device=d, dtype=torch.float64) * hi_b + 1e-6
        check(a.floor() + 1e-6, b.floor() + 1.0)
        check(a, b)


# ------------------------------------------------------------------ fused screening epilogue
@pytest.mark.parametrize("kind", ["bits", "counts"])
@pytest.mark.parametrize("n,m", [(1, 300), (257, 1000), (700, 4097)])
def test_fused_screening_scores(kind, n, m):
    """K·alpha computed inside the Gram epilogue (no Gram write) vs the unfused device path and vs a
    float64 evaluation of the oracle Gram; summation order differs -> 2e-5 relative to sum|K·alpha|."""
    from gauche_b200.screening import TanimotoScreener

    rng = np.random.default_rng(n + m)
    if kind == "bits":
        cand = (rng.random((n, 2048)) < 0.03).astype(np.float32)
        train = (rng.random((m, 2048)) < 0.03).astype(np.float32)
    else:
        cand = rng.poisson(0.05, size=(n, 2133)).astype(np.float32)
        train = rng.poisson(0.05, size=(m, 2133)).astype(np.float32)
    alpha = rng.normal(size=m).astype(np.float32)
    K = oracle.tanimoto_sim(cand, train).astype(np.float64)
    ref = K @ alpha.astype(np.float64) + 0.5
    scale = np.abs(K) @ np.abs(alpha.astype(np.float64)) + 1.0
    ct, tt, at = (torch.from_numpy(x).to(dev()) for x in (cand, train, alpha))
    fused = TanimotoScreener(tt, at, bias=0.5, block_rows=256, fused=True).scores(ct).cpu().numpy()
    tiled = TanimotoScreener(tt, at, bias=0.5, block_rows=256, fused=True, engine="mxf4x1").scores(ct).cpu().numpy()
    assert np.all(np.abs(tiled - ref) <= 2e-5 * scale)
    plain = TanimotoScreener(tt, at, bias=0.5, block_rows=256, fused=False).scores(ct).cpu().numpy()
    assert np.all(np.abs(fused - ref) <= 2e-5 * scale)
    assert np.all(np.abs(plain - ref) <= 2e-5 * scale)
    # determinism: same answer bit for bit on a second run
    again = TanimotoScreener(tt, at, bias=0.5, block_rows=256, fused=True).scores(ct).cpu().numpy()
    assert np.array_equal(fused, again)


def test_screen_topk_end_to_end():
    from gauche_b200.screening import TanimotoScreener

    g = torch.Generator().manual_seed(11)
    cand = (torch.rand((3000, 2048), generator=g) < 0.03).float().to(dev())
    train = (torch.rand((1500, 2048), generator=g) < 0.03).float().to(dev())
    alpha = torch.randn(1500, generator=g).to(dev())
    scr = TanimotoScreener(train, alpha, block_rows=1024)
    s = scr.scores(cand)
    vals, idx = scr.screen(cand, k=50, index_base=7000)
    order = torch.sort(s, descending=True, stable=True)
    assert torch.equal(idx, order.indices[:50] + 7000) and torch.equal(vals, order.values[:50])


# ------------------------------------------------------------------ packed-bit wire format
@pytest.mark.parametrize("rows,dim", [(5, 37), (64, 2048), (33, 2133), (17, 300)])
def test_packed_bit_wire_format(rows, dim):
    """A library delivered as np.packbits(..., bitorder='little') rows gives the same operands,
    norms and Gram as the dense matrix it encodes."""
    from gauche_b200 import ops
    from gauche_b200.ops import _launch_tanimoto
    from gauche_b200.packed import PackedFingerprints, pack

    rng = np.random.default_rng(rows + dim)
    x = (rng.random((rows, dim)) < 0.1).astype(np.uint8)
    wire = np.packbits(x, axis=1, bitorder="little")
    pb = PackedFingerprints.from_bits(torch.from_numpy(wire).to(dev()), dim)
    pd = pack(torch.from_numpy(x).float().to(dev()))
    assert pb.kind == "binary" and pb.kp == pd.kp
    assert torch.equal(pb.bits, pd.bits)
    assert torch.equal(pb.rowsq, pd.rowsq) and torch.equal(pb.rowsum, pd.rowsum)
    assert torch.equal(pb.q8, pd.q8)                 # lazily expanded u8 operand
    assert torch.equal(pb.q4(), pd.q4())             # lazily expanded e2m1 operand
    ref = oracle.tanimoto_sim(x.astype(np.float32), x.astype(np.float32))
    for engine in ("mxf4x1", "umma2", "popc"):
        out = torch.empty((rows, rows), dtype=torch.float32, device=dev())
        from gauche_b200 import _lib
        from gauche_b200.packed import stream_ptr
        _launch_tanimoto(_lib.load(), engine, pb, pb, out, _lib.GK_F32, 1e-6, stream_ptr(dev()))
        assert np.array_equal(out.cpu().numpy(), ref), engine


def test_screener_accepts_packed_bit_blocks():
    from gauche_b200.packed import PackedFingerprints
    from gauche_b200.screening import TanimotoScreener

    rng = np.random.default_rng(21)
    cand = (rng.random((500, 2048)) < 0.03).astype(np.uint8)
    train = (rng.random((700, 2048)) < 0.03).astype(np.uint8)
    alpha = torch.from_numpy(rng.normal(size=700).astype(np.float32)).to(dev())
    scr = TanimotoScreener(torch.from_numpy(train).float().to(dev()), alpha, block_rows=256)
    dense = scr.scores(torch.from_numpy(cand).float().to(dev()))
    wire = torch.from_numpy(np.packbits(cand, axis=1, bitorder="little")).to(dev())
    packed = scr.scores(PackedFingerprints.from_bits(wire, 2048))
    assert torch.equal(dense, packed)


# ------------------------------------------------------------------ MinMax on tensor cores
@pytest.mark.parametrize("maxval", [1, 2, 7, 31, 48, 49, 200])
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_minmax_thermometer_depths(maxval, dtype):
    """Thermometer-plane tensor-core path (depth = largest count) and the VABSDIFF4 path agree with the
    oracle bit for bit; depths beyond the limit fall back to the CUDA-core kernel."""
    from gauche_b200 import ops

    rng = np.random.default_rng(maxval)
    x1 = rng.integers(0, maxval + 1, size=(150, 300)).astype(np.float64) * (rng.random((150, 300)) < 0.2)
    x2 = rng.integers(0, maxval + 1, size=(100, 300)).astype(np.float64) * (rng.random((100, 300)) < 0.2)
    x1[0, 0] = maxval
    x2[0] = 0
    t1, t2 = torch.from_numpy(x1).to(dtype).to(dev()), torch.from_numpy(x2).to(dtype).to(dev())
    ref = oracle.minmax_sim(t1.cpu().numpy(), t2.cpu().numpy())
    d_ref, _, _ = oracle.minmax_l1(x1, x2)
    engines = ["u8", None] + (["tc"] if maxval <= ops.MINMAX_TC_MAX_PLANES else [])
    for engine in engines:
        assert np.array_equal(ops.minmax_similarity(t1, t2, engine=engine).cpu().numpy(), ref), engine
        assert np.array_equal(ops.l1_distances(t1, t2, engine=engine).cpu().numpy(), d_ref), engine
    if maxval > ops.MINMAX_TC_MAX_PLANES:
        with pytest.raises(ValueError):
            ops.minmax_similarity(t1, t2, engine="tc")


# ------------------------------------------------------------------ gradients (real-valued inputs)
@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 2e-4)])
def test_tanimoto_gradients_match_reference_autograd(dtype, tol):
    """d(sum(K*G))/dx through gk_tanimoto_bwd_coeff vs torch autograd of the reference ops (CPU fp64)."""
    import gauche_b200 as gb

    g = torch.Generator().manual_seed(5)
    x1 = torch.rand((70, 45), generator=g, dtype=torch.float64)
    x2 = torch.rand((130, 45), generator=g, dtype=torch.float64) - 0.2
    G = torch.randn((70, 130), generator=g, dtype=torch.float64)
    r1, r2 = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    (oracle.tanimoto_sim_torch(r1, r2) * G).sum().backward()
    d1 = x1.to(dtype).to(dev()).requires_grad_()
    d2 = x2.to(dtype).to(dev()).requires_grad_()
    K = gb.batch_tanimoto_sim(d1, d2)
    assert K.requires_grad
    (K * G.to(dtype).to(dev())).sum().backward()
    scale = r1.grad.abs().max().item()
    assert (d1.grad.double().cpu() - r1.grad).abs().max().item() <= tol * max(scale, 1.0)
    assert (d2.grad.double().cpu() - r2.grad).abs().max().item() <= tol * max(r2.grad.abs().max().item(), 1.0)
    # self-kernel (x1 is x2, e.g. K_zz of learnable inducing points) and only-one-side gradients
    z = x1.to(dtype).to(dev()).requires_grad_()
    gb.batch_tanimoto_sim(z, z).sum().backward()
    zr = x1.clone().requires_grad_()
    oracle.tanimoto_sim_torch(zr, zr).sum().backward()
    assert (z.grad.double().cpu() - zr.grad).abs().max().item() <= tol * max(zr.grad.abs().max().item(), 1.0)
    w = x1.to(dtype).to(dev()).requires_grad_()
    gb.TanimotoKernel().forward(w, d2.detach()).sum().backward()
    assert w.grad is not None and torch.isfinite(w.grad).all()


def test_tanimoto_gradients_batched():
    import gauche_b200 as gb

    g = torch.Generator().manual_seed(6)
    x1 = torch.rand((3, 9, 11), generator=g, dtype=torch.float64)
    x2 = torch.rand((1, 7, 11), generator=g, dtype=torch.float64)
    r1, r2 = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    oracle.tanimoto_sim_torch(r1, r2).pow(2).sum().backward()
    d1, d2 = x1.to(dev()).requires_grad_(), x2.to(dev()).requires_grad_()
    K = gb.batch_tanimoto_sim(d1, d2)
    assert K.shape == (3, 9, 7)
    K.pow(2).sum().backward()
    assert torch.allclose(d1.grad.cpu(), r1.grad, rtol=1e-9, atol=1e-10)
    assert torch.allclose(d2.grad.cpu(), r2.grad, rtol=1e-9, atol=1e-10)


@pytest.mark.parametrize("dtype,tol", [(torch.float64, 1e-9), (torch.float32, 5e-4)])
def test_minmax_gradients_match_reference_autograd(dtype, tol):
    import gauche_b200 as gb

    g = torch.Generator().manual_seed(8)
    x1 = torch.rand((70, 45), generator=g, dtype=torch.float64)
    x2 = torch.rand((130, 45), generator=g, dtype=torch.float64)
    x2[3] = x1[5]                                   # exact ties: sign(0) = 0 on both sides
    G = torch.randn((70, 130), generator=g, dtype=torch.float64)
    r1, r2 = x1.clone().requires_grad_(), x2.clone().requires_grad_()
    (oracle.minmax_sim_torch(r1, r2) * G).sum().backward()
    d1 = x1.to(dtype).to(dev()).requires_grad_()
    d2 = x2.to(dtype).to(dev()).requires_grad_()
    K = gb.batch_minmax_sim(d1, d2)
    (K * G.to(dtype).to(dev())).sum().backward()
    assert (d1.grad.double().cpu() - r1.grad).abs().max().item() <= tol * max(r1.grad.abs().max().item(), 1.0)
    assert (d2.grad.double().cpu() - r2.grad).abs().max().item() <= tol * max(r2.grad.abs().max().item(), 1.0)
    w = x1.to(dtype).to(dev()).requires_grad_()
    gb.MinMaxKernel().forward(w, d2.detach()).sum().backward()      # one-sided
    assert torch.isfinite(w.grad).all()


@pytest.mark.parametrize("n,m", [(1, 5), (130, 228), (517, 1204), (2049, 900)])
def test_engines_agree_and_fused_matches_unfused(n, m):
    """Every tensor-core engine is bit-identical to the popc engine (incl. output pitches the TMA cannot address);
    the fused K·alpha scores of each operand kind match the unfused Gram + GEMV to summation-order tolerance."""
    from gauche_b200 import ops
    from gauche_b200.screening import TanimotoScreener
    from gauche_b200.synthetic import ecfp_bits

    a = ecfp_bits(n, 2048, seed=11, device=dev())
    b = ecfp_bits(m, 2048, seed=12, device=dev())
    ref = ops.tanimoto_similarity(a, b, engine="popc")
    for eng in ("bxh", "bx", "mxf4x1", "umma2", "umma2x1"):
        got = ops.tanimoto_similarity(a, b, engine=eng)
        assert torch.equal(ref, got), (eng, n, m, float((ref - got).abs().max()))
    alpha = torch.randn(m, device=dev())
    s_u = TanimotoScreener(b, alpha, bias=0.25, fused=False).scores(a)
    for eng, kind in (("auto", 1), ("bx", 2), ("bxh", 3)):
        scr = TanimotoScreener(b, alpha, bias=0.25, fused=True, engine=eng)
        s_f = scr.scores(a)
        assert scr.last_operand_kind == kind
        assert torch.allclose(s_f, s_u, rtol=2e-5, atol=2e-5), (eng, float((s_f - s_u).abs().max()))


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.uint8])
@pytest.mark.parametrize("host_pack", ["auto", True, False])
def test_screener_scores_from_host_blocks(dtype, host_pack):
    """Dense HOST candidates (pinned or pageable, several blocks, ragged last block): host-packed or shipped
    dense, the scores are identical to the device-tensor call; count blocks always take the dense route."""
    from gauche_b200.screening import TanimotoScreener

    g = torch.Generator().manual_seed(5)
    train = (torch.rand((1500, 2048), generator=g) < 0.03).float().to(dev())
    alpha = torch.randn(1500, generator=g).to(dev())
    cand = (torch.rand((1111, 2048), generator=g) < 0.03).to(dtype)
    scr = TanimotoScreener(train, alpha, bias=0.1, block_rows=256, host_pack=host_pack, host_threads=2)
    ref = scr.scores(cand.to(dev()))
    got = scr.scores(cand.pin_memory() if dtype != torch.float64 else cand)
    assert torch.equal(ref, got)
    if host_pack is True:
        assert scr.host_bytes_copied == 1111 * (256 + 4)     # packed rows + their popcounts only
    # a block with a count in it is shipped dense and runs on the int8 kernel; the device call classifies the
    # whole tensor as counts, so the two differ only in summation order
    counts = cand.clone()
    counts[7, 9] = 3
    assert torch.allclose(scr.scores(counts), scr.scores(counts.to(dev())), rtol=1e-5, atol=1e-5)


def test_minmax_block_sparse_skips_are_exact():
    """Thermometer planes with empty blocks (incl. all-zero row tiles and a single high count): the block-sparse
    tensor-core MinMax equals the dense one and the CUDA-core kernel bit for bit, for all three output types."""
    from gauche_b200 import ops

    g = torch.Generator().manual_seed(21)
    a = (torch.rand((700, 2133), generator=g) < 0.02).float()
    b = (torch.rand((1000, 2133), generator=g) < 0.02).float()
    a[5, 100] = 9.0            # one deep count -> 9 planes, planes 2..9 nearly empty
    b[900, 100] = 7.0
    b[3, 2000] = 4.0
    a[128:256] = 0.0           # an all-zero 128-row tile of the left operand
    b[224:448] = 0.0           # an all-zero 224-row tile of the right operand
    a, b = a.to(dev()), b.to(dev())
    for dt in (torch.float32, torch.float64):
        x, y = a.to(dt), b.to(dt)
        sparse = ops.minmax_similarity(x, y, engine="tc")
        u8 = ops.minmax_similarity(x, y, engine="u8")
        dense = ops.minmax_similarity(x, y, engine="tc-dense")
        assert torch.equal(sparse, dense) and torch.equal(sparse, u8)
        for eng in ("tc-pair", "tc-single"):            # block-sparse on CTA pairs (256-row tiles) / single CTAs
            assert torch.equal(ops.minmax_similarity(x, y, engine=eng), u8), eng
            assert torch.equal(ops.minmax_similarity(x[:300], y[:77], engine=eng), u8[:300, :77]), eng
        ref = oracle.minmax_sim(x.cpu().numpy(), y.cpu().numpy())
        assert np.array_equal(sparse.cpu().numpy(), ref)
    assert torch.equal(ops.l1_distances(a, b, engine="tc"), ops.l1_distances(a, b, engine="u8"))
    assert torch.equal(ops.l1_distances(a, b, engine="tc-pair"), ops.l1_distances(a, b, engine="u8"))


@pytest.mark.parametrize("n", [1, 77, 300, 517, 1301, 5000])
def test_minmax_symmetric_gram_is_exact(n):
    """K(x, x) with x1 is x2 takes the upper-triangle kernel with mirrored stores (gk_minmax_gram_mxf4_sym): bit-identical
    to the full computation, the CUDA-core kernel and the oracle (the reference's MinMax Gram is bitwise symmetric) — incl.
    row pitches the TMA cannot address, CTA pairs and single CTAs, fp32 and raw int32 outputs."""
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_counts

    x = ecfp_counts(n, 2048, extra=85, seed=30 + n).to(dev())
    if n > 200:
        x[128:140] = 0.0
        x[7, 100] = 9.0
    for eng in ("tc", "tc-pair", "tc-single"):
        got = ops.minmax_similarity(x, x, engine=eng)
        assert ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sym", (eng, ops.last_launch)
        full = ops.minmax_similarity(x, x, engine="tc-full")
        assert ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sparse"
        assert torch.equal(got, full), (eng, n, int((got != full).sum()))
        assert torch.equal(got, got.t())
    assert torch.equal(got, ops.minmax_similarity(x, x, engine="u8"))
    if n <= 1301:
        assert np.array_equal(got.cpu().numpy(), oracle.minmax_sim(x.cpu().numpy(), x.cpu().numpy()))
    assert torch.equal(ops.l1_distances(x, x, engine="tc"), ops.l1_distances(x, x, engine="u8"))
    # a second operand object with the same values is NOT assumed symmetric
    y = x.clone()
    assert torch.equal(ops.minmax_similarity(x, y, engine="tc"), got)
    assert ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sparse"


def test_permuted_bit_operand_matches_oracle():
    """PackedFingerprints.bitsp() (gk_permute_bits_kblock, the A operand of the hybrid Gram kernel) against the numpy
    restatement, incl. widths that are not whole k-blocks and row slices."""
    import oracle.exact as ex
    from gauche_b200.packed import pack

    for n, d in ((5, 37), (130, 2048), (64, 2133), (33, 4096 + 128)):
        rng = np.random.default_rng(n + d)
        x = torch.from_numpy((rng.random((n, d)) < 0.2).astype(np.float32)).to(dev())
        p = pack(x)
        got = p.bitsp().cpu().numpy().view(np.uint32)
        ref = ex.permute_bits_kblock(p.bits.cpu().numpy().view(np.uint32)[:, :p.kp // 32])
        assert got.shape == ref.shape and np.array_equal(got, ref), (n, d)
        if n > 40:
            sl = p.rows_slice(7, 40).bitsp().cpu().numpy().view(np.uint32)
            assert np.array_equal(sl, ref[7:40])


def test_inference_mode_tensors_and_data_edits():
    """Tensors created under torch.inference_mode() have no version counter: they must work (uncached) instead of
    raising; a cached operand is re-packed after an in-place edit that bumps the version."""
    import gauche_b200 as gb
    from gauche_b200.packed import pack_cache

    rng = np.random.default_rng(3)
    x = (rng.random((70, 300)) < 0.1).astype(np.float32)
    ref = oracle.tanimoto_sim(x, x)
    with torch.inference_mode():
        t = torch.from_numpy(x).to(dev())
        assert np.array_equal(gb.batch_tanimoto_sim(t, t).cpu().numpy(), ref)
        assert np.array_equal(gb.batch_tanimoto_sim(torch.from_numpy(x), torch.from_numpy(x)).numpy(), ref)   # CPU -> GPU copy inside
    t = torch.from_numpy(x).to(dev())
    gb.batch_tanimoto_sim(t, t)
    hits = pack_cache.hits
    gb.batch_tanimoto_sim(t, t)
    assert pack_cache.hits > hits                       # same tensor, unchanged: cache hit
    t[0, :] = 1.0                                        # in-place edit bumps the version -> re-pack
    x2 = x.copy()
    x2[0, :] = 1.0
    assert np.array_equal(gb.batch_tanimoto_sim(t, t).cpu().numpy(), oracle.tanimoto_sim(x2, x2))


def test_wide_fingerprints_beyond_int32_norm_range():
    """D > 33 000 (the old hard limit of the packer): bits are exact for any width the f32 accumulators allow, real-valued
    inputs take the dense path, and count rows whose squared norm overflows int32 are routed to the dense path too."""
    from gauche_b200 import ops
    from gauche_b200.packed import pack

    rng = np.random.default_rng(9)
    d = 40_000
    a = (rng.random((9, d)) < 0.05).astype(np.float32)
    b = (rng.random((7, d)) < 0.05).astype(np.float32)
    ta, tb = torch.from_numpy(a).to(dev()), torch.from_numpy(b).to(dev())
    assert pack(ta).kind == "binary"
    for engine in ("mxf4x1", "bx", "umma2", "popc"):
        assert np.array_equal(ops.tanimoto_similarity(ta, tb, engine=engine).cpu().numpy(), oracle.tanimoto_sim(a, b)), engine
    r = rng.random((5, d)).astype(np.float64)
    tr = torch.from_numpy(r).to(dev())
    np.testing.assert_allclose(ops.tanimoto_similarity(tr, tr).cpu().numpy(), oracle.tanimoto_sim_torch(torch.from_numpy(r), torch.from_numpy(r)).numpy(), rtol=1e-10)
    c = np.full((3, d), 255.0, dtype=np.float32)         # sum of squares 2.6e9 > 2^31 - 1
    tc = torch.from_numpy(c).to(dev())
    assert pack(tc).kind == "real"
    np.testing.assert_allclose(ops.tanimoto_similarity(tc, tc).cpu().numpy(), np.ones((3, 3), np.float32), rtol=1e-5)


def test_thermometer_cache_keeps_the_deepest_expansion_and_streams_wait():
    """q4t(planes) for a shallower depth is a prefix view of the deepest expansion built so far (no second copy), and an
    operand built on one stream can be consumed on another."""
    from gauche_b200 import ops
    from gauche_b200.packed import pack

    rng = np.random.default_rng(12)
    x = rng.poisson(0.3, size=(300, 512)).astype(np.float32)
    x[0, 0] = 9
    t = torch.from_numpy(x).to(dev())
    p = pack(t)
    deep = p.q4t(9)
    shallow = p.q4t(4)
    assert shallow.data_ptr() == deep.data_ptr() and shallow.shape[1] == 4 * p.k4_bytes and shallow.stride(0) == deep.stride(0)
    assert p.q4t(9).data_ptr() == deep.data_ptr()
    ref = oracle.minmax_sim(x, x)
    side = torch.cuda.Stream()
    t2 = torch.from_numpy(x).to(dev())
    torch.cuda.synchronize()
    got_main = ops.minmax_similarity(t2, t2)             # packs + expands on the current stream (cached on t2)
    with torch.cuda.stream(side):
        got_side = ops.minmax_similarity(t2, t2)         # cache hit on another stream: must wait for the build
    side.synchronize()
    torch.cuda.synchronize()
    assert np.array_equal(got_main.cpu().numpy(), ref) and np.array_equal(got_side.cpu().numpy(), ref)
