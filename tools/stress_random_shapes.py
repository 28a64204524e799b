"""Randomised shape stress of the two kernels added late in round 2: the hybrid Tanimoto Gram (engine bxh) against the __popc
engine, and the symmetric MinMax kernel against the full computation — many small / ragged / odd shapes, bit-exact."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gauche_b200 import ops

dev = torch.device("cuda:0")
rng = np.random.default_rng(int(os.environ.get("SEED", "1")))
bad = 0
N_T = int(os.environ.get("N_TANIMOTO", "150"))
for it in range(N_T):
    n = int(rng.integers(1, 3000)) if it % 5 else int(rng.integers(1, 130))
    m = int(rng.integers(1, 3000)) if it % 7 else int(rng.integers(1, 200))
    d = int(rng.choice([1, 31, 37, 64, 100, 128, 255, 256, 257, 300, 511, 512, 700, 1024, 1152, 2048, 2133, 3000, 4096, 5000]))
    dens = float(rng.choice([0.0, 0.01, 0.05, 0.3, 0.9, 1.0]))
    a = (torch.rand((n, d), device=dev) < dens).float()
    b = a if (it % 11 == 0 and n == m) else (torch.rand((m, d), device=dev) < dens).float()
    ref = ops.tanimoto_similarity(a, b, engine="popc")
    got = ops.tanimoto_similarity(a, b, engine="bxh")
    assert ops.last_launch["tanimoto"] == "gk_tanimoto_gram_hybrid"
    if not torch.equal(ref, got):
        bad += 1
        print("TANIMOTO MISMATCH", (n, m, d, dens), int((ref != got).sum()), flush=True)
print("tanimoto hybrid: %d shapes, %d mismatches" % (N_T, bad), flush=True)
bad2 = 0
N_M = int(os.environ.get("N_MINMAX", "60"))
for it in range(N_M):
    n = int(rng.integers(1, 4000)) if it % 4 else int(rng.integers(1, 300))
    d = int(rng.choice([37, 128, 300, 1024, 2048, 2133]))
    lam = float(rng.choice([0.02, 0.3, 1.5]))
    x = torch.poisson(torch.full((n, d), lam, device=dev)).clamp_(max=40.0)
    if n > 10:
        x[3] = 0
    if float(x.max()) < 3:        # make sure the block-sparse (>= 3 planes) kernels are the ones compared
        x[0, 0] = 5.0
    for eng in ("tc-single", "tc-pair"):
        got = ops.minmax_similarity(x, x, engine=eng)
        sym = ops.last_launch["minmax"] == "gk_minmax_gram_mxf4_sym"
        full = ops.minmax_similarity(x, x, engine="tc-full")
        if not (sym and torch.equal(got, full)):
            bad2 += 1
            print("MINMAX MISMATCH", (n, d, lam, eng, sym), int((got != full).sum()), flush=True)
print("minmax symmetric: %d shapes x 2 engines, %d mismatches" % (N_M, bad2), flush=True)
sys.exit(1 if (bad or bad2) else 0)
