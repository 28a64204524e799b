"""Full-size checks (BASELINE.json configs) through size-independent properties: two independent
engines agree, row-block decomposition is exact, symmetry, unit diagonal, exact count checksums."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def test_cfg2_lipophilicity_size_full_gram_vs_oracle():
    """4200 x 4200 x 2048 (configs[1]): whole Gram bit-exact against the CPU oracle, fp32."""
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_bits

    x = ecfp_bits(4200, 2048, seed=2, adversarial=True)
    ref = oracle.tanimoto_sim(x.numpy(), x.numpy())
    xd = x.to(dev())
    for engine in ("bxh", "mxf4x1", "bx", "umma2", "umma2x1", "popc"):
        got = ops.tanimoto_similarity(xd, xd, engine=engine)
        assert np.array_equal(got.cpu().numpy(), ref), engine
        # NB: the reference Gram is NOT exactly symmetric — fl(fl(eps+n_i)+n_j) != fl(fl(eps+n_j)+n_i)
        # in the last ulp for a few thousand pairs — and the drop-in reproduces that bit for bit.
        assert float((got - got.t()).abs().max()) <= 2e-8
    d = torch.diagonal(ops.tanimoto_similarity(xd, xd))
    # like the reference, the diagonal is 1 only up to the rounding of fl(eps+n): |d-1| <= 2 ulp
    assert d[0] == 1.0 and float((d - 1.0).abs().max()) <= 2.4e-7   # zero row: eps/eps == 1


def test_cfg1_photoswitch_size_fp64():
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_counts

    x = ecfp_counts(392, 2048, extra=85, seed=1, dtype=torch.float64)   # fragprints
    xd = x.to(dev())
    assert np.array_equal(ops.tanimoto_similarity(xd, xd).cpu().numpy(), oracle.tanimoto_sim(x.numpy(), x.numpy()))
    assert np.array_equal(ops.minmax_similarity(xd, xd).cpu().numpy(), oracle.minmax_sim(x.numpy(), x.numpy()))


def test_cfg4_screening_block_properties():
    """16384 candidates x 100k train x 2048 bits: tensor-core and popc engines agree exactly on
    counts; row-block decomposition is exact; sampled sub-blocks match the CPU oracle."""
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_bits

    train = ecfp_bits(100_000, 2048, seed=4, device=dev())
    cand = ecfp_bits(16_384, 2048, seed=5, device=dev())
    cu = ops.intersection_counts(cand, train, engine="umma2")
    for other in ("bx", "mxf4x1"):
        c1 = ops.intersection_counts(cand, train, engine=other)
        assert torch.equal(cu, c1), other
        del c1
    # the hybrid kernel (default) at full size with its packed fp32 epilogue: bit-identical to the e2m1 engine's Gram
    g0 = ops.tanimoto_similarity(cand, train, engine="mxf4x1")
    g1 = ops.tanimoto_similarity(cand, train, engine="bxh")
    assert ops.last_launch["tanimoto"] == "gk_tanimoto_gram_hybrid"
    assert torch.equal(g0, g1)
    del g0, g1
    cp = ops.intersection_counts(cand, train, engine="popc")
    assert torch.equal(cu, cp)
    # checksum of checksums: sum of all intersections == sum_k colsum_cand[k] * colsum_train[k]
    exact = int((cand.double().sum(0) * train.double().sum(0)).sum().item())
    assert int(cu.sum(dtype=torch.int64).item()) == exact
    del cp
    K = ops.tanimoto_similarity(cand, train)            # default engine (mxf4x1 for bits)
    assert torch.equal(K, ops.tanimoto_similarity(cand, train, engine="umma2"))
    blk = ops.tanimoto_similarity(cand[5000:5300], train)
    assert torch.equal(K[5000:5300], blk)
    rng = np.random.default_rng(0)
    for _ in range(4):
        r0 = int(rng.integers(0, 16_384 - 256))
        c0 = int(rng.integers(0, 100_000 - 512))
        ref = oracle.tanimoto_sim(cand[r0:r0 + 256].cpu().numpy(), train[c0:c0 + 512].cpu().numpy())
        assert np.array_equal(K[r0:r0 + 256, c0:c0 + 512].cpu().numpy(), ref)
    assert float(K.min()) >= 0.0 and float(K.max()) <= 1.0


def test_cfg3_minmax_block_properties():
    """MinMax on count fingerprints, 8192 x 8192 x 2048: symmetry, unit diagonal, Tanimoto==MinMax
    on the binarised inputs (for bits min/max and and/or coincide), sampled blocks vs the oracle."""
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_counts

    x = ecfp_counts(8192, 2048, seed=3, device=dev())
    K = ops.minmax_similarity(x, x)                      # thermometer planes on the FP4 tensor cores
    assert torch.equal(K, ops.minmax_similarity(x, x, engine="u8"))   # VABSDIFF4 CUDA-core kernel
    assert torch.equal(K, K.t())
    assert torch.all(torch.diagonal(K) == 1.0)
    rng = np.random.default_rng(1)
    for _ in range(3):
        r0 = int(rng.integers(0, 8192 - 128))
        c0 = int(rng.integers(0, 8192 - 128))
        ref = oracle.minmax_sim(x[r0:r0 + 128].cpu().numpy(), x[c0:c0 + 128].cpu().numpy())
        assert np.array_equal(K[r0:r0 + 128, c0:c0 + 128].cpu().numpy(), ref)
    b = (x[:2048] > 0).float()
    d = ops.l1_distances(b, b)
    inter = ops.intersection_counts(b, b)
    n = b.sum(1).to(torch.int32)
    assert torch.equal(d, n[:, None] + n[None, :] - 2 * inter)


@pytest.mark.gpu
def test_cached_gram_is_cuda_graph_capturable():
    """GP training re-evaluates the parameter-free K(X, X) every closure: once the operands are packed (cache hit:
    no host sync), the whole call is stream-ordered and can be captured in a CUDA graph and replayed
    (tools/graph_capture_check.py: 32 -> 10 us per call at N = 392, 31 -> 25 us at N = 4200)."""
    import gauche_b200 as gb
    from gauche_b200.synthetic import ecfp_bits

    dev = torch.device("cuda:0")
    x = ecfp_bits(392, 2048, seed=2, device=dev)
    ref = gb.batch_tanimoto_sim(x, x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            gb.batch_tanimoto_sim(x, x)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = gb.batch_tanimoto_sim(x, x)
    for _ in range(3):
        out.zero_()
        graph.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)


def test_cfg3_minmax_full_size_50k():
    """BASELINE.json configs[2] at its full size: MinMax on 50 000 x 50 000 count fingerprints (10 GB fp32 Gram).
    Sampled 128 x 128 blocks (incl. the last rows/columns and diagonal blocks) bit-exact against the oracle; exact
    symmetry; unit diagonal; the VABSDIFF4 CUDA-core engine agrees on a 2048-row band."""
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_counts

    n = 50_000
    x = ecfp_counts(n, 2048, seed=3, device=dev())
    K = ops.minmax_similarity(x, x)                      # block-sparse thermometer planes on the FP4 tensor cores
    assert ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sym"    # upper-triangle tiles + mirrored stores
    assert K.shape == (n, n) and K.dtype == torch.float32
    for r in range(0, n, 5000):                          # every element was written, and written symmetrically
        assert torch.equal(K[r:r + 5000], K[:, r:r + 5000].t()), r
    assert torch.all(torch.diagonal(K) == 1.0)
    rng = np.random.default_rng(3)
    corners = [(0, 0), (n - 128, n - 128), (n - 128, 0), (12_345, 12_345)]
    corners += [(int(rng.integers(0, n - 128)), int(rng.integers(0, n - 128))) for _ in range(6)]
    for r0, c0 in corners:
        ref = oracle.minmax_sim(x[r0:r0 + 128].cpu().numpy(), x[c0:c0 + 128].cpu().numpy())
        assert np.array_equal(K[r0:r0 + 128, c0:c0 + 128].cpu().numpy(), ref), (r0, c0)
        assert torch.equal(K[r0:r0 + 128, c0:c0 + 128], K[c0:c0 + 128, r0:r0 + 128].t())     # symmetry, block-wise
    band = ops.minmax_similarity(x[20_000:22_048], x, engine="u8")
    assert torch.equal(K[20_000:22_048], band)
    assert float(K.min()) >= 0.0 and float(K.max()) <= 1.0


def test_cfg5_sparse_gp_shard_and_inducing_gram():
    """BASELINE.json configs[4], one rank's share at 8 GPUs: K_xz = K(X[125 000 shard], Z[4096]) and K_zz = K(Z, Z),
    Z = the first 4096 rows of the library (base(X,Z), base(Z,Z), base(X,X,diag=True)).  Sampled
    256 x 512 blocks vs the oracle, exact intersection counts vs the __popc engine, the shard equals the same rows
    of a second shard-boundary decomposition, diag mode."""
    import gauche_b200 as gb
    from gauche_b200 import ops
    from gauche_b200.screening import shard_bounds
    from gauche_b200.synthetic import ecfp_bits

    start, stop = shard_bounds(1_000_000, 8, 3)
    assert stop - start == 125_000
    lib = ecfp_bits(4096 + 125_000, 2048, seed=6, device=dev())     # frozen inducing rows + this rank's shard
    z, xs = lib[:4096], lib[4096:]
    kern = gb.TanimotoKernel()
    kxz = kern.forward(xs, z)
    kzz = kern.forward(z, z)
    assert kxz.shape == (125_000, 4096) and kzz.shape == (4096, 4096)
    rng = np.random.default_rng(5)
    blocks = [(0, 0), (125_000 - 256, 4096 - 512)] + [(int(rng.integers(0, 125_000 - 256)), int(rng.integers(0, 4096 - 512)))
                                                     for _ in range(4)]
    for r0, c0 in blocks:
        ref = oracle.tanimoto_sim(xs[r0:r0 + 256].cpu().numpy(), z[c0:c0 + 512].cpu().numpy())
        assert np.array_equal(kxz[r0:r0 + 256, c0:c0 + 512].cpu().numpy(), ref), (r0, c0)
    assert np.array_equal(kzz.cpu().numpy(), oracle.tanimoto_sim(z.cpu().numpy(), z.cpu().numpy()))
    cnt = ops.intersection_counts(xs, z)
    assert torch.equal(cnt, ops.intersection_counts(xs, z, engine="popc"))
    del cnt
    # row-shard decomposition is exact: the same rows computed as part of another block boundary
    assert torch.equal(kxz[60_000:61_000], kern.forward(xs[60_000:61_000], z))
    d = kern.forward(xs[:1000], xs[:1000], diag=True)
    assert d.shape == (1000,) and torch.all(d == 1)


def test_fused_scores_vs_fp64_truth():
    """The fused screening epilogue uses num * rcp.approx(den) inside the K·alpha sum.  Against an fp64 evaluation of
    the same sum, its error must be of the same order as that of the reference's own fp32 computation (oracle torch
    port on the CPU: correctly rounded fp32 Gram, fp32 matmul) — i.e. the approximate reciprocal is not what limits
    the accuracy of a 10^5-term fp32 sum."""
    from gauche_b200.screening import TanimotoScreener
    from gauche_b200.synthetic import ecfp_bits

    m, n = 20_000, 512
    train = ecfp_bits(m, 2048, seed=4)
    cand = ecfp_bits(n, 2048, seed=5)
    alpha = torch.randn(m, generator=torch.Generator().manual_seed(6)) * 1e-3
    truth = torch.from_numpy(oracle.tanimoto_sim(cand.double().numpy(), train.double().numpy())) @ alpha.double()
    ref32 = (oracle.tanimoto_sim_torch(cand, train) @ alpha).double()          # the reference's fp32 arithmetic
    scale = float(truth.abs().max())
    for fused in (True, False):
        got = TanimotoScreener(train.to(dev()), alpha.to(dev()), fused=fused).scores(cand.to(dev())).cpu().double()
        err_gpu = float((got - truth).abs().max()) / scale
        err_ref = float((ref32 - truth).abs().max()) / scale
        assert err_gpu <= 1e-5, (fused, err_gpu)
        assert err_gpu <= 4.0 * err_ref + 1e-7, (fused, err_gpu, err_ref)


def test_random_shape_stress_of_hybrid_and_symmetric_kernels():
    """Many small / ragged / odd shapes (rows 1..3000, widths 1..5000 bits, densities 0..1): the hybrid Tanimoto kernel equals
    the __popc engine and the symmetric MinMax kernel equals the full computation, bit for bit (tools/stress_random_shapes.py
    runs the same loop with thousands of shapes)."""
    from gauche_b200 import ops

    rng = np.random.default_rng(7)
    for it in range(80):
        n = int(rng.integers(1, 3000)) if it % 5 else int(rng.integers(1, 130))
        m = int(rng.integers(1, 3000)) if it % 7 else int(rng.integers(1, 200))
        d = int(rng.choice([1, 31, 37, 128, 255, 256, 257, 511, 700, 1152, 2048, 2133, 4096, 5000]))
        dens = float(rng.choice([0.0, 0.01, 0.05, 0.3, 1.0]))
        a = (torch.rand((n, d), device=dev()) < dens).float()
        b = (torch.rand((m, d), device=dev()) < dens).float()
        got = ops.tanimoto_similarity(a, b, engine="bxh")
        assert ops.last_launch["tanimoto"] == "gk_tanimoto_gram_hybrid"
        assert torch.equal(got, ops.tanimoto_similarity(a, b, engine="popc")), (n, m, d, dens)
    for it in range(24):
        n = int(rng.integers(1, 4000)) if it % 4 else int(rng.integers(1, 300))
        d = int(rng.choice([37, 300, 2048, 2133]))
        x = torch.poisson(torch.full((n, d), float(rng.choice([0.02, 0.3, 1.5])), device=dev())).clamp_(max=40.0)
        x[0, 0] = 5.0                                    # >= 3 thermometer planes: the block-sparse kernels
        full = ops.minmax_similarity(x, x, engine="tc-full")
        for eng in ("tc-single", "tc-pair"):
            got = ops.minmax_similarity(x, x, engine=eng)
            assert ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sym"
            assert torch.equal(got, full), (n, d, eng)
