"""Measure the practical int8 tensor peak on this B200 (cuBLASLt via torch._int_mm, 8192^3) as burst
and sustained figures, and our Gram kernel sustained, each with nvidia-smi clocks/power sampled under load."""
import os, subprocess, sys, tempfile, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gauche_b200 import _lib
from gauche_b200.ops import _launch_tanimoto
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.synthetic import ecfp_bits

dev = torch.device("cuda:0")

class Smi:
    def __init__(self):
        self.f = tempfile.NamedTemporaryFile("w+", delete=False)
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                                   "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=self.f)
    def stop(self):
        self.p.terminate(); self.p.wait(); self.f.flush(); self.f.seek(0)
        rows = [l.split(",") for l in self.f.read().splitlines() if l.count(",") >= 4]
        clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
        hot = [c for c, p in zip(clk, pw) if p > 0.6 * max(pw)]
        return {"samples": len(rows), "sm_mhz_median_loaded": statistics.median(hot) if hot else None,
                "sm_mhz_min_loaded": min(hot) if hot else None,
                "power_w_max": max(pw) if pw else None, "power_cap_active": any("Active" == r[2].strip() for r in rows),
                "hw_slowdown": any("Active" == r[3].strip() for r in rows), "sw_thermal": any("Active" == r[4].strip() for r in rows)}

def timed(fn, seconds):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # burst: best of 10
    best = 1e9
    for _ in range(10):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    smi = Smi(); time.sleep(0.3)
    n = max(10, int(seconds * 1e3 / best))
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) / n
    return best, sustained, smi.stop()

N = 8192
a = torch.randint(-3, 4, (N, N), dtype=torch.int8, device=dev)
b = torch.randint(-3, 4, (N, N), dtype=torch.int8, device=dev).t()   # column-major B as cuBLASLt prefers
burst, sus, clk = timed(lambda: torch._int_mm(a, b), 4.0)
print("int_mm 8192^3: burst %.1f TOP/s, sustained %.1f TOP/s" % (2 * N**3 / burst / 1e9, 2 * N**3 / sus / 1e9), clk)
x = torch.randn((N, N), dtype=torch.bfloat16, device=dev); y = torch.randn((N, N), dtype=torch.bfloat16, device=dev)
burst, sus, clk = timed(lambda: torch.matmul(x, y), 4.0)
print("bf16 8192^3:   burst %.1f TFLOP/s, sustained %.1f TFLOP/s" % (2 * N**3 / burst / 1e9, 2 * N**3 / sus / 1e9), clk)
del a, b, x, y
lib = _lib.load()
R, M = 16384, 100000
train = pack(ecfp_bits(M, 2048, seed=4, device=dev)); cand = pack(ecfp_bits(R, 2048, seed=5, device=dev))
out = torch.empty((R, M), dtype=torch.float32, device=dev)
for eng in sys.argv[1:] or ["umma2x1", "umma2"]:
    burst, sus, clk = timed(lambda: _launch_tanimoto(lib, eng, cand, train, out, _lib.GK_F32, 1e-6, stream_ptr(dev)), 4.0)
    print("gram %-8s: burst %.1f Gpairs/s (%.1f TOP/s), sustained %.1f Gpairs/s (%.1f TOP/s)" % (
        eng, R * M / burst / 1e6, 4096 * R * M / burst / 1e9, R * M / sus / 1e6, 4096 * R * M / sus / 1e9), clk)
