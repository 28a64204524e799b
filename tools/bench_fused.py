"""Time the fused screening kernels (K·alpha in the epilogue, no Gram write) on resident packed operands."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200.packed import pack
from gauche_b200.screening import TanimotoScreener
from gauche_b200.synthetic import ecfp_bits
dev = torch.device("cuda:0")
R, M = 16384, 100000
train = pack(ecfp_bits(M, 2048, seed=4, device=dev)); cand = pack(ecfp_bits(R, 2048, seed=5, device=dev))
alpha = torch.randn(M, device=dev) * 1e-3
ref = None
for eng in ["mxf4as", "mxf4x1"]:
    scr = TanimotoScreener(train, alpha, block_rows=R, engine=eng)
    out = torch.empty(R, dtype=torch.float32, device=dev)
    scr.scores(cand, out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): scr.scores(cand, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    if ref is None: ref = out.clone()
    print("fused %-7s: %.3f ms -> %.1f Gpairs/s ; max|diff vs first| %.3g" % (eng, ms, R * M / ms / 1e6, float((out - ref).abs().max())))
