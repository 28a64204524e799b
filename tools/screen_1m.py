"""configs[3] at full size: 100 000 train x 1 000 000 candidate 2048-bit fingerprints, candidate rows sharded over
the ranks (torchrun, one process per GPU), fused Tanimoto·alpha scores + local top-k + one NCCL all-gather.
Checks: identical top-k on every rank; sampled candidates' scores against a float64 evaluation of the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gauche_b200.packed import pack
from gauche_b200.screening import TanimotoScreener, shard_bounds
from gauche_b200.synthetic import ecfp_bits

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
N_CAND, N_TRAIN, K = int(os.environ.get("N_CAND", 1_000_000)), 100_000, 1000
start, stop = shard_bounds(N_CAND, world, rank)
train_dense = ecfp_bits(N_TRAIN, 2048, seed=4, device=dev)
alpha = (torch.randn(N_TRAIN, generator=torch.Generator().manual_seed(6)) * 1e-3).to(dev)
scr = TanimotoScreener(pack(train_dense), alpha, bias=0.1, block_rows=16384)
# this rank's shard, generated in chunks and kept only as packed operands (1 KB / molecule)
# (global chunk g = rows [65536 g, 65536 (g+1)) is always generated from seed 1000 + g, so the library is the
#  same for every world size and the global top-k can be compared across runs)
CH = 65536
chunks = []
for g in range(start // CH, (stop + CH - 1) // CH):
    g0, g1 = g * CH, min(N_CAND, (g + 1) * CH)
    dense = ecfp_bits(g1 - g0, 2048, seed=1000 + g, device=dev)
    lo, hi = max(start, g0) - g0, min(stop, g1) - g0
    chunks.append(pack(dense[lo:hi].contiguous()))
    chunks[-1].q4()
    del dense
torch.cuda.synchronize()
if world > 1: dist.barrier()
def run():
    scores = torch.empty(stop - start, dtype=torch.float32, device=dev)
    off = 0
    for p in chunks:
        scr.scores(p, out=scores[off:off + p.rows]); off += p.rows
    from gauche_b200.screening import gather_topk, topk
    v, i = topk(scores, K, index_base=start)
    return scores, gather_topk(v, i, K)
run(); torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter(); scores, (vals, idx) = run(); torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
if world > 1:
    all_idx = [torch.empty_like(idx) for _ in range(world)]; dist.all_gather(all_idx, idx)
    assert all(torch.equal(a, idx) for a in all_idx), "ranks disagree on the global top-k"
# sampled exactness: candidates of the first local chunk vs the oracle in float64
import oracle
g = start // CH
sample = ecfp_bits(min(N_CAND, (g + 1) * CH) - g * CH, 2048, seed=1000 + g, device=dev)[start - g * CH:start - g * CH + 64].cpu().numpy()
K64 = oracle.tanimoto_sim(sample, train_dense.cpu().numpy()).astype(np.float64)
ref = K64 @ alpha.cpu().numpy().astype(np.float64) + 0.1
err = np.abs(scores[:64].cpu().numpy() - ref).max()
if rank == 0:
    print("screened %d candidates x %d train on %d GPU(s) in %.2f ms (%.0f Gpairs/s); top-1 score %.5f idx %d; "
          "max |score - oracle| on 64 samples %.2e; top-k index checksum %d" % (
              N_CAND, N_TRAIN, world, dt * 1e3, N_CAND * N_TRAIN / dt / 1e9, float(vals[0]), int(idx[0]), err,
              int((idx * torch.arange(1, K + 1, device=dev)).sum().item())), flush=True)
assert err < 1e-4
if world > 1: dist.destroy_process_group()
