"""cfg3: MinMax Gram on count fingerprints (50k x 50k x 2048, streamed in row blocks) + the CUDA-core
Tanimoto engines for comparison.  Prints Gpairs/s per engine (CUDA events, resident packed operands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import _lib
from gauche_b200.ops import _launch_minmax, _launch_tanimoto
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.synthetic import ecfp_bits, ecfp_counts

dev = torch.device("cuda:0")
lib = _lib.load()
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
R = 8192
x = pack(ecfp_counts(N, 2048, seed=3, device=dev))
xb = pack(ecfp_bits(N, 2048, seed=3, device=dev))
out = torch.empty((R, N), dtype=torch.float32, device=dev)
st = stream_ptr(dev)

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

blk = x.rows_slice(0, R); blkb = xb.rows_slice(0, R)
ms = timeit(lambda: _launch_minmax(lib, blk, x, out, _lib.GK_F32, 1e-6, st, "tc"))
print("minmax tc (thermometer, %d planes): %.2f ms per %dx%d block -> %.1f Gpairs/s ; full %dx%d would take %.1f ms" % (max(x.maxval, 1), ms, R, N, R * N / ms / 1e6, N, N, ms * N / R))
ms = timeit(lambda: _launch_minmax(lib, blk, x, out, _lib.GK_F32, 1e-6, st, "u8"))
print("minmax u8 (VABSDIFF4)   : %.2f ms per %dx%d block -> %.1f Gpairs/s ; full %dx%d would take %.1f ms" % (ms, R, N, R * N / ms / 1e6, N, N, ms * N / R))
ms = timeit(lambda: _launch_tanimoto(lib, "u8", blk, x, out, _lib.GK_F32, 1e-6, st))
print("tanimoto u8 (dp4a)      : %.2f ms -> %.1f Gpairs/s" % (ms, R * N / ms / 1e6))
ms = timeit(lambda: _launch_tanimoto(lib, "popc", blkb, xb, out, _lib.GK_F32, 1e-6, st))
print("tanimoto popc           : %.2f ms -> %.1f Gpairs/s" % (ms, R * N / ms / 1e6))
ms = timeit(lambda: _launch_tanimoto(lib, "umma2", blk, x, out, _lib.GK_F32, 1e-6, st))
print("tanimoto umma2 (counts) : %.2f ms -> %.1f Gpairs/s" % (ms, R * N / ms / 1e6))
