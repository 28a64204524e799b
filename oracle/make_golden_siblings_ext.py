"""Generate tests/golden/sibling_vectors_ext.npz: the UNMODIFIED reference sibling kernels
(/root/reference/gauche/kernels/fingerprint_kernels/*_kernel.py, stub gpytorch) on the parts of their surface beyond
2-D integer inputs — leading batch dimensions (incl. broadcasting), real-valued inputs, and ``last_dim_is_batch``
through the kernel classes' ``covar_dist``.  Cases the reference itself rejects are recorded as ``raises``."""
from __future__ import annotations

import importlib.util
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import OUT, REF, counts_like, load_reference  # noqa: E402
from siblings import NAMES  # noqa: E402


def class_name(name: str) -> str:
    return "".join(p.capitalize() for p in name.split("_")) + "Kernel"


def main():
    load_reference()   # installs the stub gpytorch module
    torch.set_num_threads(1)
    fns, classes = {}, {}
    for name in NAMES:
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + "_kernel.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        fns[name] = getattr(mod, f"batch_{name}_sim")
        classes[name] = getattr(mod, class_name(name))
    rng = np.random.default_rng(20261021)

    def bits(*shape, p=0.2, dt=np.float32):
        return (rng.random(shape) < p).astype(dt)

    inputs = {
        # (x1, x2, last_dim_is_batch)
        "batch_sq_f32": (bits(3, 16, 96), bits(3, 16, 96), False),
        "batch_sq_f64": (bits(2, 12, 40, dt=np.float64), bits(2, 12, 40, dt=np.float64), False),
        "batch_bcast_f32": (bits(3, 16, 96), bits(16, 96), False),
        "batch4d_f32": (bits(2, 1, 9, 70), bits(1, 3, 9, 70), False),
        "batch_rect_f32": (bits(2, 10, 64), bits(2, 7, 64), False),
        "batch_counts_f64": (np.stack([counts_like(rng, 12), counts_like(rng, 12)]).astype(np.float64),
                             np.stack([counts_like(rng, 12), counts_like(rng, 12)]).astype(np.float64), False),
        "real_sq_f64": (rng.random((24, 50)) * 2.0, rng.random((24, 50)) * 2.0, False),
        "real_rect_f32": ((rng.random((20, 33)) * 3.0).astype(np.float32), (rng.random((12, 33)) * 3.0).astype(np.float32), False),
        "real_batch_f32": ((rng.random((2, 8, 21))).astype(np.float32), (rng.random((2, 8, 21))).astype(np.float32), False),
        "ldb_bits_f32": (bits(6, 5, p=0.5), bits(6, 5, p=0.5), True),
        "ldb_rect_f64": (bits(7, 4, p=0.5, dt=np.float64), bits(3, 4, p=0.5, dt=np.float64), True),
    }
    arrays, manifest = {}, []
    for key, (x1, x2, ldb) in inputs.items():
        arrays[f"in_{key}_x1"], arrays[f"in_{key}_x2"] = x1, x2
        for name in NAMES:
            t1, t2 = torch.from_numpy(x1), torch.from_numpy(x2)
            try:
                if ldb:
                    out = classes[name]().covar_dist(t1, t2, last_dim_is_batch=True).numpy()
                else:
                    out = fns[name](t1, t2).numpy()
                arrays[f"out_{key}_{name}"] = out
                manifest.append({"input": key, "kernel": name, "last_dim_is_batch": ldb, "out_dtype": str(out.dtype),
                                 "out_shape": list(out.shape)})
            except Exception as exc:   # the reference raises (shape quirk / non-binary): recorded, not an error
                manifest.append({"input": key, "kernel": name, "last_dim_is_batch": ldb, "raises": type(exc).__name__})
    np.savez_compressed(os.path.join(OUT, "sibling_vectors_ext.npz"), **arrays)
    json.dump({"generator": "oracle/make_golden_siblings_ext.py", "torch": torch.__version__, "cases": manifest},
              open(os.path.join(OUT, "sibling_vectors_ext.json"), "w"), indent=1)
    print(len(manifest), "cases;", sum(1 for c in manifest if "raises" in c), "raise in the reference")
    for c in manifest:
        if "raises" in c:
            print("  raises:", c["input"], c["kernel"], c["raises"])


if __name__ == "__main__":
    main()
