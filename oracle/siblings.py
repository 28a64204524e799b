"""Exact-integer numpy restatement of GAUCHE's twelve sibling fingerprint kernels (test infrastructure only).

Follows gauche/kernels/fingerprint_kernels/{dice,braun_blanquet,faith,forbes,inner_product,intersection,otsuka,
rand,rogers_tanimoto,russell_rao,sogenfrei,sokal_sneath}_kernel.py line by line, INCLUDING their shape quirks
(last-row indexing `x[-1]`, untransposed `x2_norm`), with the contraction done in exact integers and the float
operations replayed one at a time in the reference's order and dtype.  2-D integer-valued inputs.
Pinned bit-for-bit against tests/golden/sibling_vectors.npz (oracle/make_golden_siblings.py).
"""
from __future__ import annotations

import numpy as np

NAMES = ["dice", "braun_blanquet", "faith", "forbes", "inner_product", "intersection", "otsuka", "rand",
         "rogers_tanimoto", "russell_rao", "sogenfrei", "sokal_sneath"]
TO_DOUBLE = {"braun_blanquet", "forbes", "intersection", "otsuka", "rand", "rogers_tanimoto", "russell_rao",
             "sogenfrei", "sokal_sneath"}


def sibling_sim(name: str, x1: np.ndarray, x2: np.ndarray, eps: float = 1e-6) -> np.ndarray:
    x1, x2 = np.asarray(x1), np.asarray(x2)
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    if x1.ndim != 2 or x2.ndim != 2:
        raise NotImplementedError("2-D inputs only")
    if name == "intersection" and (np.any(x1 > 1) or np.any(x2 > 1)):
        raise ValueError("Tensors must be binary-valued")
    dt = x1.dtype.type if x1.dtype.kind == "f" else np.float32
    e = dt(eps)
    a, b = x1.astype(np.float64), x2.astype(np.float64)
    dot = np.rint(a @ b.T).astype(np.int64).astype(dt)          # exact, then the tensor dtype
    n1 = np.rint(a.sum(-1)).astype(np.int64).astype(dt)[:, None]
    n2 = np.rint(b.sum(-1)).astype(np.int64).astype(dt)[:, None]  # [M,1], NOT transposed
    n = x1.shape[-1]
    d = dt(np.sum((x1[-1] == 0) & (x2[-1] == 0)))                # last rows only
    two, three = dt(2), dt(3)
    if name == "dice":
        q = (two * dot + e) / ((n1 + n2.T) + e)
    elif name == "braun_blanquet":
        q = (dot + e) / (np.maximum(n1[-1], n2[-1]) + e)
    elif name == "faith":
        q = ((two * dot + d) + e) / dt(2 * n + eps)
    elif name == "forbes":
        q = (dt(n) * dot + e) / ((n1 + n2.T) + e)
    elif name == "inner_product":
        q = dot + e
    elif name == "intersection":
        q = (dot + d) + e
    elif name == "otsuka":
        q = (dot + e) / np.sqrt((n1 + n2.T) + e)
    elif name == "rand":
        q = ((dot + d) + e) / dt(n + eps)
    elif name == "rogers_tanimoto":
        q = ((dot + d) + e) / ((((two * n1 + two * n2) - three * dot) + d) + e)
    elif name == "russell_rao":
        q = (dot + e) / dt(n + eps)
    elif name == "sogenfrei":
        t = dot + e
        q = (t * t) / ((n1 + n2.T) + e)
    elif name == "sokal_sneath":
        q = (dot + e) / (((two * n1 + two * n2) - three * dot) + e)
    else:
        raise KeyError(name)
    q = np.asarray(q, dtype=dt)
    if name in TO_DOUBLE:
        q = q.astype(np.float64)
    return np.where(q < 0, q.dtype.type(0), q)
