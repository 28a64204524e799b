"""Small fixed launch sequence for ncu: N launches of the Tanimoto Gram kernel on resident operands."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import _lib
from gauche_b200.ops import _launch_tanimoto
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.synthetic import ecfp_bits

p = argparse.ArgumentParser()
p.add_argument("--rows", type=int, default=4096)
p.add_argument("--train", type=int, default=25600)
p.add_argument("--launches", type=int, default=4)
p.add_argument("--engine", default="bx")
p.add_argument("--fused", action="store_true", help="fused K·alpha screening launch instead of the Gram")
p.add_argument("--minmax", action="store_true", help="MinMax Gram of --rows count fingerprints (block-sparse thermometer planes)")
p.add_argument("--dtype", default="f32")
p.add_argument("--bits", type=int, default=2048)
a = p.parse_args()
dev = torch.device("cuda:0")
lib = _lib.load()
if a.minmax:
    from gauche_b200 import ops
    from gauche_b200.synthetic import ecfp_counts
    x = ecfp_counts(a.rows, a.bits, seed=3, device=dev)
    ops.minmax_similarity(x, x)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
    ev[0].record()
    for i in range(a.launches):
        k = ops.minmax_similarity(x, x)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
    print("minmax ms per launch:", ["%.3f" % m for m in ms], "Gpairs/s (last): %.1f" % (a.rows * a.rows / ms[-1] / 1e6))
    sys.exit(0)
train = pack(ecfp_bits(a.train, a.bits, seed=4, device=dev))
cand = pack(ecfp_bits(a.rows, a.bits, seed=5, device=dev))
odt = torch.float32 if a.dtype == "f32" else torch.float64
code = _lib.GK_F32 if a.dtype == "f32" else _lib.GK_F64
out = torch.empty((a.rows, a.train), dtype=odt, device=dev)
scr = None
if a.fused:
    from gauche_b200.screening import TanimotoScreener
    scr = TanimotoScreener(train, torch.randn(a.train, device=dev) * 1e-3, block_rows=a.rows,
                           engine="auto" if a.engine == "bx" else a.engine)
    sc = torch.empty((a.rows,), dtype=torch.float32, device=dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.launches + 1)]
ev[0].record()
for i in range(a.launches):
    if scr is not None:
        scr.scores(cand, out=sc)
    else:
        _launch_tanimoto(lib, a.engine, cand, train, out, code, 1e-6, stream_ptr(dev))
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.launches)]
print("ms per launch:", ["%.3f" % m for m in ms], "Gpairs/s (last): %.1f" % (a.rows * a.train / ms[-1] / 1e6))
