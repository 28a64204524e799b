"""A/B of the host packer (gk_host_pack_bits) builds: dense fp32 0/1 blocks [16384 x 2048] -> wire format, 8 pinned
blocks in rotation, threads in {4, 8, 16, all}.  Usage: python tools/host_pack_bench.py lib1.so [lib2.so ...]"""
import ctypes, os, sys, time
import torch

R, D = 16384, 2048
blocks = [((torch.rand((R, D)) < 0.03).float()).pin_memory() for _ in range(8)]
out = torch.empty((R * (D // 32 + 1),), dtype=torch.int32).pin_memory()
flag = ctypes.c_int32(0)
for path in sys.argv[1:]:
    lib = ctypes.CDLL(os.path.abspath(path))
    fn = lib.gk_host_pack_bits
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                   ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.c_int]
    for threads in (4, 8, 16, os.cpu_count()):
        for rep in range(2):
            t0 = time.perf_counter()
            for b in blocks * 2:
                fn(b.data_ptr(), 0, R, D, D, out.data_ptr(), D // 32, out.data_ptr() + R * (D // 32) * 4, ctypes.byref(flag), threads)
            dt = (time.perf_counter() - t0) / 16
        print(f"{os.path.basename(path):16s} threads {threads:3d}: {dt * 1e3:.3f} ms per block  {R * D * 4 / dt / 1e9:.1f} GB/s  not_binary={flag.value}", flush=True)
