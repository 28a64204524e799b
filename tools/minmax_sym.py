"""cfg3 (MinMax K(x, x) on 50 000 count fingerprints) through the public call: the symmetric kernel (upper-triangle tiles +
mirrored stores) vs the full computation; also smaller sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import _lib
if os.environ.get("GAUCHE_B200_LIB"):
    _lib.LIB_PATH = os.environ["GAUCHE_B200_LIB"]
from gauche_b200 import ops
from gauche_b200.synthetic import ecfp_counts

dev = torch.device("cuda:0")
for n in [int(v) for v in os.environ.get("SIZES", "4200,8192,20000,50000").split(",")]:
    x = ecfp_counts(n, 2048, seed=3, device=dev)
    out = {}
    for eng in ("tc", "tc-full", "tc-single", "tc-pair"):
        k = ops.minmax_similarity(x, x, engine=eng)
        k2 = ops.minmax_similarity(x, x, engine=eng)      # two result buffers in the caching allocator before the timed loop
        torch.cuda.synchronize()
        del k2
        name = ops.last_launch["minmax"]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            k = ops.minmax_similarity(x, x, engine=eng)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[eng] = k[::7, ::5].clone()
        print(f"n={n} engine {eng:9s} {name:28s}: {ms:.3f} ms  {n * n / ms / 1e6:.0f} Gpairs/s", flush=True)
        del k
    print("   sampled results identical:", bool(torch.equal(out["tc"], out["tc-full"]) and torch.equal(out["tc"], out["tc-single"]) and torch.equal(out["tc"], out["tc-pair"])), flush=True)
