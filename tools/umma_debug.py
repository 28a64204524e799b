"""Diagnostic: compare the tcgen05 Gram kernel with the popc kernel on growing problems and
describe the mismatch pattern (used while bringing the kernel up on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import ops

ENGINE = sys.argv[1] if len(sys.argv) > 1 else "umma"
dev = torch.device("cuda:0")
torch.manual_seed(0)
ok_all = True
for (n, m, d) in [(128, 256, 128), (128, 256, 256), (128, 256, 2048), (256, 512, 2048), (100, 300, 2048),
                  (1000, 3000, 2048), (4200, 4200, 2048)]:
    x1 = (torch.rand((n, d), device=dev) < 0.1).float()
    x2 = (torch.rand((m, d), device=dev) < 0.1).float()
    ref = ops.intersection_counts(x1, x2, engine="popc")
    torch.cuda.synchronize()
    try:
        got = ops.intersection_counts(x1, x2, engine=ENGINE)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"[{n}x{m}x{d}] umma raised: {e}")
        ok_all = False
        break
    bad = (got != ref)
    nb = int(bad.sum())
    print(f"[{n}x{m}x{d}] mismatches: {nb} / {n*m}")
    if nb:
        ok_all = False
        idx = bad.nonzero()[:10].tolist()
        for (i, j) in idx:
            print(f"   ({i},{j}) got {int(got[i,j])} ref {int(ref[i,j])}")
        rows_bad = bad.any(1).nonzero().flatten()[:20].tolist()
        cols_bad = bad.any(0).nonzero().flatten()[:20].tolist()
        print("   bad rows:", rows_bad, " bad cols:", cols_bad)
        print("   got sum", int(got.sum()), "ref sum", int(ref.sum()))
        break
    # float epilogue too
    a = ops.tanimoto_similarity(x1, x2, engine=ENGINE)
    b = ops.tanimoto_similarity(x1, x2, engine="popc")
    print(f"   float epilogue equal: {bool(torch.equal(a, b))}")
    if not torch.equal(a, b): ok_all = False
    a = ops.tanimoto_similarity(x1.double(), x2.double(), engine=ENGINE)
    b = ops.tanimoto_similarity(x1.double(), x2.double(), engine="popc")
    print(f"   fp64 epilogue equal: {bool(torch.equal(a, b))}")
    if not torch.equal(a, b): ok_all = False
print("UMMA_DEBUG", ENGINE, "OK" if ok_all else "FAIL")
