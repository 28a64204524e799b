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
    for engine in ("bx", "mxf4x1", "umma2", "umma2x1", "popc"):
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
    cp = ops.intersection_counts(cand, train, engine="popc")
    assert torch.equal(cu, cp)
    # checksum of checksums: sum of all intersections == sum_k colsum_cand[k] * colsum_train[k]
    exact = int((cand.double().sum(0) * train.double().sum(0)).sum().item())
    assert int(cu.sum(dtype=torch.int64).item()) == exact
    del cp
    K = ops.tanimoto_similarity(cand, train)            # default engine (bx for bits)
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
