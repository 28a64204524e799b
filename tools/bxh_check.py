"""Check of the hybrid Gram kernel (engine "bxh": A = permuted packed bits expanded into tensor memory, B = ready-made
e2m1 by TMA): exactness against the popc engine on a few shapes, then kernel timings against the e2m1 engine on the cfg4
block (Gram to HBM and fused K.alpha).  GAUCHE_B200_LIB=<path> loads an alternative build of the library."""
import os, sys
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
shapes = [(128, 192, 256), (128, 192, 512), (130, 228, 2048), (517, 1204, 2048), (300, 520, 1024 + 128), (2049, 900, 4096),
          (64, 64, 37), (1, 1, 2048), (1000, 3, 300)]
for (n, m, d) in [] if os.environ.get("BX_CHECK_SKIP_EXACT") else shapes:
    g = torch.Generator().manual_seed(n + m + d)
    a = (torch.rand((n, d), generator=g) < 0.3).float().to(dev)
    b = (torch.rand((m, d), generator=g) < 0.3).float().to(dev)
    try:
        kh = ops.tanimoto_similarity(a, b, engine="bxh")
        torch.cuda.synchronize()
        assert ops.last_launch["tanimoto"] == "gk_tanimoto_gram_hybrid", ops.last_launch
        kp = ops.tanimoto_similarity(a, b, engine="popc")
        eq = bool(torch.equal(kh, kp))
        print((n, m, d), "bxh float Gram equal:", eq, "wrong:", int((kh != kp).sum()), "max|diff|",
              float((kh - kp).abs().max()), flush=True)
        ok &= eq
    except Exception as e:
        print("bxh FAILED", (n, m, d), repr(e)[:300], flush=True)
        ok = False
        break
print("EXACT" if ok else "MISMATCH", flush=True)
if not ok:
    sys.exit(1)

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


ref = None
for rnd in range(2):
    for eng in ("bxh", "mxf4x1", "bx"):
        try:
            ms = timeit(lambda: ops._launch_tanimoto(lib, eng, cand, train, out, _lib.GK_F32, 1e-6, st))
            print(f"gram {eng:7s}: {ms:.3f} ms  {R * M / ms / 1e6:.0f} Gpairs/s", flush=True)
            if rnd == 0:
                cs = out[::97, ::89].double().sum().item()
                if ref is None:
                    ref = cs
                print("   sampled checksum", cs, "equal to first:", cs == ref, flush=True)
        except Exception as e:
            print("gram", eng, "FAILED", repr(e)[:300], flush=True)
if os.environ.get("BX_CHECK_F64"):          # fp64 output (the reference's default dtype for GP work): 8 B per pair
    del out
    out64 = torch.empty((R, M), dtype=torch.float64, device=dev)
    for eng in ("bxh", "mxf4x1", "bxh", "mxf4x1"):
        ms = timeit(lambda: ops._launch_tanimoto(lib, eng, cand, train, out64, _lib.GK_F64, 1e-6, st))
        print(f"gram fp64 {eng:7s}: {ms:.3f} ms  {R * M / ms / 1e6:.0f} Gpairs/s", flush=True)
    del out64
    out = torch.empty((R, M), dtype=torch.float32, device=dev)
if os.environ.get("BX_CHECK_SUSTAIN"):      # ~2 s of back-to-back launches: the power-capped regime
    for eng in ("bxh", "mxf4x1"):
        ms = timeit(lambda: ops._launch_tanimoto(lib, eng, cand, train, out, _lib.GK_F32, 1e-6, st), reps=1300)
        print(f"gram {eng:7s} sustained (1300 launches): {ms:.3f} ms  {R * M / ms / 1e6:.0f} Gpairs/s", flush=True)
ref = None
for eng in (("bxh", "auto") if not os.environ.get("BX_CHECK_GRAM_ONLY") else ()):
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

# ---- other sizes: where does the hybrid kernel start to pay?
for (r, m) in () if os.environ.get("BX_CHECK_NO_SIZES") else ((392, 392), (4200, 4200), (4096, 4096), (8192, 8192), (16384, 16384), (125000, 4096), (50000, 50000)):
    a = pack(ecfp_bits(r, 2048, seed=7, device=dev))
    b = a if r == m else pack(ecfp_bits(m, 2048, seed=8, device=dev))
    o = torch.empty((r, m), dtype=torch.float32, device=dev)
    line = f"{r} x {m}:"
    for eng in ("bxh", "mxf4x1"):
        ms = timeit(lambda: ops._launch_tanimoto(lib, eng, a, b, o, _lib.GK_F32, 1e-6, st), reps=50 if r * m < 1e8 else 10)
        line += f"  {eng} {ms * 1e3:.1f} us ({r * m / ms / 1e6:.0f} Gpairs/s)"
    print(line, flush=True)
    del o
