"""First-light check of the bit-expanding Gram kernel (engine "bx"): exactness against the popc engine on a few
shapes, then kernel timings of bx vs the ready-made-e2m1 engine on the cfg4 block.  GAUCHE_B200_LIB=<path> loads an
alternative build of the library (e.g. csrc/libgauche_b200_alt.so)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gauche_b200 import _lib
if os.environ.get("GAUCHE_B200_LIB"):
    _lib.LIB_PATH = os.environ["GAUCHE_B200_LIB"]
from gauche_b200 import ops
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.screening import TanimotoScreener
from gauche_b200.synthetic import ecfp_bits

dev = torch.device("cuda:0")
print("lib:", _lib.LIB_PATH, flush=True)
ok = True
for (n, m, d) in [] if os.environ.get("BX_CHECK_SKIP_EXACT") else [(128, 224, 256), (128, 224, 512), (130, 228, 2048), (517, 1204, 2048), (300, 520, 1024 + 128), (2049, 900, 4096), (64, 64, 37)]:
    g = torch.Generator().manual_seed(n + m + d)
    a = (torch.rand((n, d), generator=g) < 0.3).float().to(dev)
    b = (torch.rand((m, d), generator=g) < 0.3).float().to(dev)
    ref = ops.intersection_counts(a, b, engine="popc")
    try:
        got = ops.intersection_counts(a, b, engine="bx")
        torch.cuda.synchronize()
        kt = ops.tanimoto_similarity(a, b, engine="bxh")
        torch.cuda.synchronize()
        kp = ops.tanimoto_similarity(a, b, engine="popc")
        print("   bxh float Gram equal:", bool(torch.equal(kt, kp)), "wrong:", int((kt != kp).sum()), "max|diff|", float((kt - kp).abs().max()), flush=True)
    except Exception as e:
        print("bx FAILED", (n, m, d), repr(e)[:300], flush=True)
        ok = False
        break
    bad = (ref != got)
    print((n, m, d), "counts equal:", not bool(bad.any()), "wrong:", int(bad.sum()), "of", bad.numel(),
          "max|diff|", int((ref - got).abs().max()), flush=True)
    if bad.any():
        ok = False
        idx = bad.nonzero()[:5]
        for i, j in idx.tolist():
            print("   ", (i, j), "ref", int(ref[i, j]), "got", int(got[i, j]))
    kf = ops.tanimoto_similarity(a, b, engine="bx")
    kr = ops.tanimoto_similarity(a, b, engine="popc")
    print("   float Gram equal:", bool(torch.equal(kf, kr)), flush=True)
print("EXACT" if ok else "MISMATCH", flush=True)

# ---- timings on the cfg4 block
R, M = 16384, 100000
train = pack(ecfp_bits(M, 2048, seed=4, device=dev))
cand = pack(ecfp_bits(R, 2048, seed=5, device=dev))
out = torch.empty((R, M), dtype=torch.float32, device=dev)
lib = _lib.load()
st = stream_ptr(dev)
alpha = (torch.randn(M, generator=torch.Generator().manual_seed(6)) * 1e-3).to(dev)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for eng in ("bxh", "bx", "mxf4x1", "umma2"):
    try:
        ms = timeit(lambda: ops._launch_tanimoto(lib, eng, cand, train, out, _lib.GK_F32, 1e-6, st))
        print(f"gram {eng:7s}: {ms:.3f} ms  {R * M / ms / 1e6:.0f} Gpairs/s", flush=True)
    except Exception as e:
        print("gram", eng, "FAILED", repr(e)[:300], flush=True)
ref = None
for eng in (("auto", "bx") if not os.environ.get("BX_CHECK_GRAM_ONLY") else ()):
    try:
        scr = TanimotoScreener(train, alpha, engine=eng)
        s = scr.scores(cand)
        ms = timeit(lambda: scr.scores(cand))
        print(f"fused scores {eng:7s} (kind {scr.last_operand_kind}): {ms:.3f} ms  {R * M / ms / 1e6:.0f} Gpairs/s", flush=True)
        if ref is None:
            ref = s
        else:
            print("   max |diff| vs first:", float((s - ref).abs().max()), flush=True)
    except Exception as e:
        print("fused", eng, "FAILED", repr(e)[:300], flush=True)
