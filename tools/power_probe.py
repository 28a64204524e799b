"""Is the Gram kernel's rate data dependent (power)?  Times the tensor-core engines on real ECFP-like bits, on all-zero
fingerprints and on dense random bits, with nvidia-smi power / clock samples during each run."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gauche_b200 import _lib, ops
from gauche_b200.packed import pack, stream_ptr
from gauche_b200.synthetic import ecfp_bits

dev = torch.device("cuda:0")
lib = _lib.load()
st = stream_ptr(dev)
R, M = 16384, 100000
out = torch.empty((R, M), dtype=torch.float32, device=dev)


def sample_start():
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "20", "-i", "0"],
                         stdout=f, stderr=subprocess.DEVNULL)
    return p, f


def sample_stop(p, f):
    p.terminate(); p.wait()
    f.flush(); f.seek(0)
    rows = [l.split(",") for l in f.read().splitlines() if "," in l]
    os.unlink(f.name)
    clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
    top = sorted(zip(pw, clk))[-max(1, len(pw) // 3):]
    return sum(c for _, c in top) / len(top), max(pw) if pw else 0


def run(name, a, b, eng, reps):
    for _ in range(3):
        ops._launch_tanimoto(lib, eng, a, b, out, _lib.GK_F32, 1e-6, st)
    torch.cuda.synchronize()
    p, f = sample_start()
    time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops._launch_tanimoto(lib, eng, a, b, out, _lib.GK_F32, 1e-6, st)
    e1.record()
    torch.cuda.synchronize()
    clk, pw = sample_stop(p, f)
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:10s} {eng:7s} reps {reps:4d}: {ms:.3f} ms  {R * M / ms / 1e6:6.0f} Gpairs/s   sm clock (top-power samples) {clk:.0f} MHz, max power {pw:.0f} W", flush=True)


data = {
    "ecfp": (ecfp_bits(R, 2048, seed=5, device=dev), ecfp_bits(M, 2048, seed=4, device=dev)),
    "zeros": (torch.zeros((R, 2048), device=dev), torch.zeros((M, 2048), device=dev)),
    "dense50": ((torch.rand((R, 2048), device=dev) < 0.5).float(), (torch.rand((M, 2048), device=dev) < 0.5).float()),
}
for name, (xa, xb) in data.items():
    a, b = pack(xa), pack(xb)
    for eng in ("mxf4x1", "bx", "umma2"):
        for reps in (20, 300):
            run(name, a, b, eng, reps)
    del a, b
