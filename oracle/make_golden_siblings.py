"""Generate tests/golden/sibling_vectors.npz by running the UNMODIFIED reference sibling kernels
(/root/reference/gauche/kernels/fingerprint_kernels/*_kernel.py, stub gpytorch) in the build container."""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, counts_like, ecfp_like, load_reference  # noqa: E402
from siblings import NAMES  # noqa: E402


def main():
    load_reference()   # installs the stub gpytorch module
    torch.set_num_threads(1)
    fns = {}
    for name in NAMES:
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + "_kernel.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        fns[name] = getattr(mod, f"batch_{name}_sim")
    rng = np.random.default_rng(20261020)
    inputs = {
        "kat": (np.array([[1, 0, 1, 1]], np.float64), np.array([[1, 1, 0, 0]], np.float64)),   # the reference tests' KAT
        "ones": (np.ones((2, 2), np.float64), np.ones((2, 2), np.float64)),
        "bits_sq_f32": (ecfp_like(rng, 64).astype(np.float32), ecfp_like(rng, 64).astype(np.float32)),
        "bits_sq_f64": (ecfp_like(rng, 48).astype(np.float64), ecfp_like(rng, 48).astype(np.float64)),
        "bits_rect_f32": (ecfp_like(rng, 40).astype(np.float32), ecfp_like(rng, 24).astype(np.float32)),
        "bits_small_f64": ((rng.random((9, 37)) < 0.3).astype(np.float64), (rng.random((9, 37)) < 0.3).astype(np.float64)),
        "counts_sq_f64": (counts_like(rng, 32).astype(np.float64), counts_like(rng, 32).astype(np.float64)),
        "counts_rect_f32": (counts_like(rng, 20).astype(np.float32), counts_like(rng, 12).astype(np.float32)),
    }
    inputs["bits_sq_f32"][0][0] = 0     # zero row
    arrays, manifest = {}, []
    for key, (x1, x2) in inputs.items():
        arrays[f"in_{key}_x1"], arrays[f"in_{key}_x2"] = x1, x2
        for name in NAMES:
            try:
                out = fns[name](torch.from_numpy(x1), torch.from_numpy(x2)).numpy()
                arrays[f"out_{key}_{name}"] = out
                manifest.append({"input": key, "kernel": name, "out_dtype": str(out.dtype), "out_shape": list(out.shape)})
            except Exception as exc:   # reference raises (shape quirk / non-binary): recorded, not an error
                manifest.append({"input": key, "kernel": name, "raises": type(exc).__name__})
    np.savez_compressed(os.path.join(OUT, "sibling_vectors.npz"), **arrays)
    json.dump({"generator": "oracle/make_golden_siblings.py", "torch": torch.__version__, "cases": manifest},
              open(os.path.join(OUT, "sibling_vectors.json"), "w"), indent=1)
    print(len(manifest), "cases;", sum(1 for c in manifest if "raises" in c), "raise in the reference")


if __name__ == "__main__":
    main()
