"""Can the cached small-Gram path be captured in a CUDA graph (GP training loops re-evaluate K(X,X) every closure)?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gauche_b200 as gb
from gauche_b200.synthetic import ecfp_bits

dev = torch.device("cuda:0")
for n in (392, 4200):
    x = ecfp_bits(n, 2048, seed=2, device=dev)
    ref = gb.batch_tanimoto_sim(x, x)            # packs + caches the operands
    torch.cuda.synchronize()
    def timed(fn, reps=200):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e6
    eager_us = timed(lambda: gb.batch_tanimoto_sim(x, x))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): gb.batch_tanimoto_sim(x, x)
    torch.cuda.current_stream().wait_stream(s)
    try:
        with torch.cuda.graph(g):
            out = gb.batch_tanimoto_sim(x, x)
        g.replay(); torch.cuda.synchronize()
        ok = torch.equal(out, ref)
        graph_us = timed(g.replay)
        print(f"N={n}: eager {eager_us:.1f} us/call, graph replay {graph_us:.1f} us/call, identical={ok}")
    except Exception as e:  # noqa
        print(f"N={n}: capture failed: {type(e).__name__}: {str(e)[:300]}")
