import sys, torch
sys.path.insert(0, ".")
from gauche_b200 import ops
from gauche_b200.synthetic import ecfp_counts
dev = torch.device("cuda:0")
for n in (8192, 50000):
    x = ecfp_counts(n, 2048, seed=3, device=dev)
    ref = None
    for eng in ("tc-single", "tc-pair"):
        k = ops.minmax_similarity(x, x, engine=eng)
        if ref is None: ref = k
        else: print("  equal to single:", bool(torch.equal(ref, k)))
        del k
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            k = ops.minmax_similarity(x, x, engine=eng)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"n={n} {eng}: {ms:.3f} ms  {n*n/ms/1e6:.0f} Gpairs/s", flush=True)
    del ref
