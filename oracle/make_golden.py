"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python oracle/make_golden.py            # needs /root/reference; not run on the GPU box

The reference modules import gpytorch at module top (tanimoto_kernel.py:6-8, minmax_kernel.py:7),
which is not installed here, so a stub ``gpytorch.kernels.Kernel`` (a bare torch.nn.Module) is
injected before loading the two files with importlib — the functions under test are pure torch.
Inputs are seeded; outputs are whatever the reference returns, stored verbatim.
"""
from __future__ import annotations

import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/gauche/kernels/fingerprint_kernels"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    gp, k = types.ModuleType("gpytorch"), types.ModuleType("gpytorch.kernels")

    class Kernel(torch.nn.Module):
        def __init__(self, **kw):
            super().__init__()

    k.Kernel = Kernel
    gp.kernels = k
    sys.modules["gpytorch"], sys.modules["gpytorch.kernels"] = gp, k
    mods = {}
    for name in ("tanimoto_kernel", "minmax_kernel"):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["tanimoto_kernel"], mods["minmax_kernel"]


def ecfp_like(rng, rows, dim=2048, mean=48.0, sd=14.0):
    """Seeded ECFP-like bit matrix: row popcount ~ clip(N(48,14),8,160), Zipf-ish column weights."""
    w = 1.0 / (1.0 + np.arange(dim)) ** 0.6
    w = w[rng.permutation(dim)]
    w /= w.sum()
    x = np.zeros((rows, dim), dtype=np.uint8)
    for r in range(rows):
        n = int(np.clip(round(rng.normal(mean, sd)), 8, min(160, dim)))
        x[r, rng.choice(dim, size=n, replace=False, p=w)] = 1
    return x


def counts_like(rng, rows, dim=2048, extra=85):
    x = ecfp_like(rng, rows, dim).astype(np.int64)
    x = x * (1 + rng.poisson(0.3, size=x.shape))
    frag = rng.poisson(0.5, size=(rows, extra))
    return np.clip(np.concatenate([x, frag], axis=1), 0, 255)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)  # deterministic reduction order for the real-valued cases
    tan, mm = load_reference()
    rng = np.random.default_rng(20261019)
    os.makedirs(OUT, exist_ok=True)
    arrays, manifest = {}, []

    def add(fn_name, fn, x1, x2, note, **kw):
        cid = f"c{len(manifest):02d}"
        t1 = torch.from_numpy(np.ascontiguousarray(x1))
        t2 = t1 if x2 is None else torch.from_numpy(np.ascontiguousarray(x2))
        out = fn(t1, t2, **kw).numpy()
        arrays[cid + "_x1"] = x1
        if x2 is not None:
            arrays[cid + "_x2"] = x2
        arrays[cid + "_out"] = out
        manifest.append({"id": cid, "fn": fn_name, "self": x2 is None, "note": note,
                         "kwargs": {k: v for k, v in kw.items()},
                         "in_dtype": str(x1.dtype), "out_dtype": str(out.dtype),
                         "out_shape": list(out.shape)})

    T, M = tan.batch_tanimoto_sim, mm.batch_minmax_sim
    tk, mk = tan.TanimotoKernel(), mm.MinMaxKernel()

    # --- the reference's own known-answer inputs (test_tanimoto_kernel.py:18-79) ---
    for scale in (1.0, 2.0, 10.0, 10000.0, 0.1):
        x = (scale * np.ones((2, 2))).astype(np.float32)
        add("batch_tanimoto_sim", T, x, x.copy(), f"KAT equal inputs x{scale}")
    add("batch_tanimoto_sim", T, np.ones((2, 2), np.float32), 2 * np.ones((2, 2), np.float32), "KAT 2/3")
    a = np.array([[1, 3], [2, 4]], np.float64)
    b = np.array([[4, 2], [3, 1]], np.float64)
    add("batch_tanimoto_sim", T, a, b, "KAT very unequal fp64")
    add("batch_tanimoto_sim", T, a[None], b[None], "KAT batch dim fp64")

    # --- bit fingerprints, incl. adversarial rows ---
    for dt in (np.float32, np.float64):
        x1 = ecfp_like(rng, 96)
        x2 = ecfp_like(rng, 80)
        x1[0] = 0           # all-zero row
        x1[1] = 1           # all-one row
        x1[2] = 0; x1[2, 77] = 1   # single bit
        x1[3] = x1[4]       # duplicate
        x2[0] = 0
        x2[5] = x1[4]       # shared molecule across operands
        add("batch_tanimoto_sim", T, x1.astype(dt), x2.astype(dt), f"ecfp bits cross {np.dtype(dt).name}")
        add("batch_tanimoto_sim", T, x1.astype(dt), None, f"ecfp bits self {np.dtype(dt).name}")
    # D not a multiple of 32/128, N or M = 1
    xs = (rng.random((7, 37)) < 0.3).astype(np.float32)
    add("batch_tanimoto_sim", T, xs, xs[:1].copy(), "bits D=37, M=1")
    add("batch_tanimoto_sim", T, xs[:1].copy(), xs, "bits D=37, N=1")
    # int64 input -> float32 output (test_tanimoto_kernel.py:88 uses randint)
    xi = rng.integers(0, 2, size=(10, 5)).astype(np.int64)
    add("batch_tanimoto_sim", T, xi, None, "int64 randint bits")
    # count fingerprints (fragprints 2048+85)
    for dt in (np.float32, np.float64):
        c1 = counts_like(rng, 64).astype(dt)
        c2 = counts_like(rng, 48).astype(dt)
        c1[0] = 0
        add("batch_tanimoto_sim", T, c1, c2, f"fragprint counts {np.dtype(dt).name}")
    # batch broadcasting
    xb = (rng.random((2, 10, 5)) < 0.5).astype(np.float32)
    add("batch_tanimoto_sim", T, xb, None, "batch [2,10,5] self")
    b1 = (rng.random((3, 1, 7, 33)) < 0.4).astype(np.float64)
    b2 = (rng.random((1, 2, 5, 33)) < 0.4).astype(np.float64)
    add("batch_tanimoto_sim", T, b1, b2, "broadcast [3,1,7,33]x[1,2,5,33]")
    # real-valued
    for dt in (np.float32, np.float64):
        r1 = rng.random((40, 100)).astype(dt)
        r2 = rng.random((24, 100)).astype(dt)
        add("batch_tanimoto_sim", T, r1, r2, f"real-valued {np.dtype(dt).name}")
    rn = rng.normal(size=(9, 16)).astype(np.float64)
    add("batch_tanimoto_sim", T, rn, None, "real-valued with negatives fp64 (clamp path)")
    # kernel-class surface: last_dim_is_batch and diag
    xl = (rng.random((6, 4)) < 0.5).astype(np.float32)
    add("TanimotoKernel.covar_dist", lambda p, q, **kw: tk.covar_dist(p, q, **kw), xl, xl.copy(),
        "last_dim_is_batch", last_dim_is_batch=True)
    add("TanimotoKernel.forward", lambda p, q, **kw: tk.forward(p, q, **kw), xl, xl.copy(), "diag", diag=True)
    add("TanimotoKernel.forward", lambda p, q, **kw: tk.forward(p, q, **kw), xb, xb.copy(), "diag batch", diag=True)

    # --- MinMax (test_minmax_kernel.py:17-38) ---
    add("batch_minmax_sim", M, np.ones((2, 2), np.float64), np.ones((2, 2), np.float64), "KAT ones")
    add("batch_minmax_sim", M, np.array([[1, 0, 1, 1]], np.float64), np.array([[1, 1, 0, 0]], np.float64), "KAT 1/4")
    for dt in (np.float32, np.float64):
        c1 = counts_like(rng, 64).astype(dt)
        c2 = counts_like(rng, 48).astype(dt)
        c1[0] = 0
        c2[0] = 0
        add("batch_minmax_sim", M, c1, c2, f"fragprint counts {np.dtype(dt).name}")
        add("batch_minmax_sim", M, c1, None, f"fragprint counts self {np.dtype(dt).name}")
        x1 = ecfp_like(rng, 40).astype(dt)
        x2 = ecfp_like(rng, 33).astype(dt)
        add("batch_minmax_sim", M, x1, x2, f"ecfp bits {np.dtype(dt).name}")
        r1 = rng.random((40, 100)).astype(dt)
        r2 = rng.random((24, 100)).astype(dt)
        add("batch_minmax_sim", M, r1, r2, f"real-valued {np.dtype(dt).name}")
    mb = rng.poisson(0.7, size=(2, 10, 5)).astype(np.float64)
    add("batch_minmax_sim", M, mb, None, "batch [2,10,5] self")
    add("MinMaxKernel.covar_dist", lambda p, q, **kw: mk.covar_dist(p, q, **kw), xl.astype(np.float64),
        xl.astype(np.float64), "last_dim_is_batch", last_dim_is_batch=True)
    add("MinMaxKernel.forward", lambda p, q, **kw: mk.forward(p, q, **kw), xl, xl.copy(), "diag", diag=True)

    np.savez_compressed(os.path.join(OUT, "reference_vectors.npz"), **arrays)
    meta = {"generator": "oracle/make_golden.py", "reference": "leojklarner/gauche v0.1.6 (/root/reference)",
            "torch": torch.__version__, "threads": 1, "cases": manifest}
    with open(os.path.join(OUT, "reference_vectors.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print(f"wrote {len(manifest)} cases, {sum(a.nbytes for a in arrays.values())/1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
