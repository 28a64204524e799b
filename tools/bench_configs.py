"""Timings of the five BASELINE.json configs through the PUBLIC drop-in API (dense tensors in, kernel
matrix out), next to the reference algorithm (oracle/torch_port.py: same ATen ops) on the same GPU and,
for the small configs, on the host CPU.  Wall clock with synchronisation, median of several calls."""
import os, statistics, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauche_b200 as gb
from gauche_b200.packed import pack_cache
from gauche_b200.synthetic import ecfp_bits, ecfp_counts
from oracle.torch_port import minmax_sim_torch, tanimoto_sim_torch

dev = torch.device("cuda:0")

def wall(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3

def line(name, pairs, ours_ms, ref_gpu_ms=None, ref_cpu_ms=None, note=""):
    s = "%-46s ours %9.3f ms (%8.1f Gpairs/s)" % (name, ours_ms, pairs / ours_ms / 1e6)
    if ref_gpu_ms: s += " | torch-on-B200 %9.3f ms (x%.1f)" % (ref_gpu_ms, ref_gpu_ms / ours_ms)
    if ref_cpu_ms: s += " | torch-CPU %9.1f ms (x%.0f)" % (ref_cpu_ms, ref_cpu_ms / ours_ms)
    print(s + ("  " + note if note else ""), flush=True)

torch.set_num_threads(os.cpu_count() or 1)
print("host threads:", torch.get_num_threads())
# cfg1: Photoswitch-size fragprints, fp64, full Gram (392 x 392 x 2133)
x = ecfp_counts(392, 2048, extra=85, seed=1, dtype=torch.float64); xd = x.to(dev)
for cache in (False, True):
    pack_cache.enabled = cache; pack_cache.clear()
    line("cfg1 tanimoto 392^2 fp64 fragprints cache=%s" % cache, 392 * 392, wall(lambda: gb.batch_tanimoto_sim(xd, xd)),
         wall(lambda: tanimoto_sim_torch(xd, xd)), wall(lambda: tanimoto_sim_torch(x, x), reps=3, warm=1))
    line("cfg1 minmax   392^2 fp64 fragprints cache=%s" % cache, 392 * 392, wall(lambda: gb.batch_minmax_sim(xd, xd)),
         wall(lambda: minmax_sim_torch(xd, xd)), wall(lambda: minmax_sim_torch(x, x), reps=3, warm=1))
# cfg2: Lipophilicity-size bits, fp32, full Gram (4200 x 4200 x 2048) and train-size 3360
x = ecfp_bits(4200, 2048, seed=2); xd = x.to(dev)
for cache in (False, True):
    pack_cache.enabled = cache; pack_cache.clear()
    line("cfg2 tanimoto 4200^2 fp32 bits cache=%s" % cache, 4200 * 4200, wall(lambda: gb.batch_tanimoto_sim(xd, xd)),
         wall(lambda: tanimoto_sim_torch(xd, xd)), wall(lambda: tanimoto_sim_torch(x, x), reps=3, warm=1))
pack_cache.enabled = True
# cfg3: MinMax on count ECFP, 50k (streamed in 8192-row blocks by the caller)
n3 = 50000
c = ecfp_counts(n3, 2048, seed=3, device=dev)
def cfg3():
    gb.batch_minmax_sim(c, c)       # K(x, x): the symmetric kernel (upper-triangle tiles, mirrored stores); 10 GB result
line("cfg3 minmax 50k x 50k fp32 counts", n3 * n3, wall(cfg3, reps=3, warm=2),
     note="(reference cdist on B200 for one 2048 x 50k slab: %.1f ms)" % wall(lambda: minmax_sim_torch(c[:2048], c), reps=3, warm=1))
del c
# cfg4: screening block 16384 x 100k bits fp32 through the kernel API
tr = ecfp_bits(100000, 2048, seed=4, device=dev); ca = ecfp_bits(16384, 2048, seed=5, device=dev)
line("cfg4 tanimoto 16384 x 100k fp32 bits", 16384 * 100000, wall(lambda: gb.batch_tanimoto_sim(ca, tr), reps=5),
     wall(lambda: tanimoto_sim_torch(ca, tr), reps=3, warm=1))
# vendor int8 baseline: cuBLASLt int8 GEMM (torch._int_mm) + the reference's elementwise epilogue in torch
ca8, tr8 = ca.to(torch.int8), tr.to(torch.int8).t().contiguous().t()
na_, nb_ = ca.sum(1, keepdim=True), tr.sum(1, keepdim=True).t()
def int_mm_path():
    dot = torch._int_mm(ca8, tr8.t() if tr8.shape[0] == ca8.shape[1] else tr8.t()).float()
    return ((dot + 1e-6) / (1e-6 + na_ + nb_ - dot)).clamp_min_(0)
try:
    tr8c = tr.to(torch.int8).t().contiguous()      # [2048, 100000] column operand of _int_mm
    def int_mm_path():
        dot = torch._int_mm(ca8, tr8c).float()
        return ((dot + 1e-6) / (1e-6 + na_ + nb_ - dot)).clamp_min_(0)
    ms = wall(int_mm_path, reps=3, warm=1)
    print("cfg4 vendor baseline torch._int_mm (cuBLASLt int8) + torch epilogue: %.3f ms (%.1f Gpairs/s)" % (ms, 16384 * 100000 / ms / 1e6), flush=True)
    del tr8c
except Exception as exc:
    print("cfg4 vendor baseline torch._int_mm unavailable:", exc)
# cfg5: sparse-GP blocks: K_xz (125k x 4096 per GPU at 8 GPUs) and K_zz (4096^2)
z = tr[:4096].contiguous(); xs = ecfp_bits(125000, 2048, seed=6, device=dev)
line("cfg5 K_xz 125k x 4096 fp32 bits (1/8 shard)", 125000 * 4096, wall(lambda: gb.batch_tanimoto_sim(xs, z), reps=5),
     wall(lambda: tanimoto_sim_torch(xs, z), reps=3, warm=1))
line("cfg5 K_zz 4096^2 fp32 bits", 4096 * 4096, wall(lambda: gb.batch_tanimoto_sim(z, z), reps=7), wall(lambda: tanimoto_sim_torch(z, z)))
