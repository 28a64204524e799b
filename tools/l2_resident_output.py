"""Does the Gram store stream hurt through HBM write-back or through the SM<->L2 fabric?
Same kernel, same operands (L2 resident), 592 tiles = 4 full waves; the output either overwrites ONE 68 MB
buffer every launch (dirty lines can stay in the 126 MB L2) or rotates over 12 buffers (816 MB: every
launch's lines must be written back to HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import _lib
from gauche_b200.ops import _launch_tanimoto
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.synthetic import ecfp_bits

dev = torch.device("cuda:0")
lib = _lib.load()
engine = sys.argv[1] if len(sys.argv) > 1 else "mxf4x1"
rows, cols = 37 * 128, (16 * 224 if engine.startswith("mxf4") else 14 * 256)
cand = pack(ecfp_bits(rows, 2048, seed=5, device=dev))
train = pack(ecfp_bits(cols, 2048, seed=4, device=dev))
for nbuf in (1, 12):
    outs = [torch.empty((rows, cols), dtype=torch.float32, device=dev) for _ in range(nbuf)]
    for rep in range(2):
        n = 48
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            _launch_tanimoto(lib, engine, cand, train, outs[i % nbuf], _lib.GK_F32, 1e-6, stream_ptr(dev))
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print("%s  %d output buffer(s) of %.0f MB: %.2f us/launch  %.1f Gpairs/s" %
          (engine, nbuf, rows * cols * 4 / 1e6, ms * 1e3, rows * cols / ms / 1e6))
