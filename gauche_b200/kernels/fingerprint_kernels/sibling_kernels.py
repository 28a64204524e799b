"""The other twelve fingerprint kernels of ``gauche.kernels.fingerprint_kernels`` on the B200 path.

Each is an epilogue over the exact integer statistics the tensor-core Gram kernel produces —
``dot[i,j]`` (tcgen05 intersection counts / u8 dot products), the L1 row norms, ``d`` = common zeros of the
LAST rows and the feature count — evaluated by ``gk_sibling_epilogue`` in the reference's operation order,
so bit / count fingerprints give bit-identical results.  The reference's shape quirks are reproduced, not
fixed (last-row ``d`` and Braun-Blanquet denominator, untransposed ``x2_norm`` in Rogers-Tanimoto and
Sokal-Sneath which needs N == M).  Reference: gauche/kernels/fingerprint_kernels/<name>_kernel.py.

Integer-valued (bit or count <= 255) 2-D inputs only; real-valued or batched inputs raise
``NotImplementedError`` (there is no torch fallback in this package).
"""
from __future__ import annotations

import torch

from ... import _lib
from ..._kernel_base import Kernel
from ...ops import _check_inputs, _needs_grad, _require_cuda, _similarity
from ...packed import pack_cache, pack, stream_ptr

_IDS = {name: i for i, name in enumerate([
    "dice", "braun_blanquet", "faith", "forbes", "inner_product", "intersection", "otsuka", "rand",
    "rogers_tanimoto", "russell_rao", "sogenfrei", "sokal_sneath"])}
# kernels whose reference ends with `.to(dtype=torch.double)`
_TO_DOUBLE = {"braun_blanquet", "forbes", "intersection", "otsuka", "rand", "rogers_tanimoto", "russell_rao",
              "sogenfrei", "sokal_sneath"}
_NEEDS_D = {"faith", "intersection", "rand", "rogers_tanimoto"}


def _sibling_sim(name: str, x1: torch.Tensor, x2: torch.Tensor, eps: float) -> torch.Tensor:
    _check_inputs(x1, x2)
    if _needs_grad(x1, x2):
        raise NotImplementedError(f"gauche_b200: {name} kernel is not differentiable w.r.t. its inputs")
    if x1.ndim != 2 or x2.ndim != 2:
        raise NotImplementedError(f"gauche_b200: {name} kernel supports 2-D inputs (the reference's last-row "
                                  "indexing is ill-defined for batches)")
    if name == "intersection" and (bool(torch.any(x1 > 1)) or bool(torch.any(x2 > 1))):
        raise ValueError("Tensors must be binary-valued")        # intersection_kernel.py:36-37
    in_device = x1.device
    same = x2 is x1
    if not x1.is_cuda:
        dev = _require_cuda()
        x1 = x1.to(dev)
        x2 = x1 if same else x2.to(dev)
    dev = x1.device
    n, m, dim = x1.shape[0], x2.shape[0], x1.shape[1]
    if name in ("rogers_tanimoto", "sokal_sneath") and not (n == m or m == 1):
        raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b ({m}) at non-singleton "
                           "dimension 0")                         # x1_norm + x2_norm (untransposed) in the reference
    compute = x1.dtype if x1.dtype in (torch.float32, torch.float64) else torch.float32
    out_dtype = torch.float64 if name in _TO_DOUBLE else compute
    out = torch.empty((n, m), dtype=out_dtype, device=dev)
    if n == 0 or m == 0:
        return out.to(in_device)
    p1 = pack_cache.get(x1) if pack_cache.enabled else pack(x1)
    p2 = p1 if same else (pack_cache.get(x2) if pack_cache.enabled else pack(x2))
    if "real" in (p1.kind, p2.kind):
        raise NotImplementedError(f"gauche_b200: {name} kernel runs on bit / count fingerprints (integers in [0,255])")
    dot = _similarity("tanimoto", x1, x2, 0.0, None, raw_counts=True)        # exact int32 <x1_i, x2_j>
    d = None
    if name in _NEEDS_D:
        d = ((x1[-1] == 0) & (x2[-1] == 0)).sum().to(torch.int64).reshape(1)  # last rows only, like the reference
    codes = {torch.float32: _lib.GK_F32, torch.float64: _lib.GK_F64}
    with torch.cuda.device(dev):
        rc = _lib.load().gk_sibling_epilogue(
            _IDS[name], dot.data_ptr(), dot.stride(0) if n > 1 else max(m, 1), n, m,
            p1.rowsum.data_ptr(), p2.rowsum.data_ptr(), d.data_ptr() if d is not None else None, dim,
            codes[compute], out.data_ptr(), out.stride(0) if n > 1 else max(m, 1), codes[out_dtype], eps,
            stream_ptr(dev))
    _lib.check(rc, "gk_sibling_epilogue")
    return out if out.device == in_device else out.to(in_device)


def _make(name: str, class_name: str):
    def batch_sim(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
        if x1.ndim < 2 or x2.ndim < 2:
            raise ValueError("Tensors must have a batch dimension")
        return _sibling_sim(name, x1, x2, eps)

    batch_sim.__name__ = f"batch_{name}_sim"
    batch_sim.__doc__ = (f"{name} similarity over the last two dims (reference "
                         f"gauche/kernels/fingerprint_kernels/{name}_kernel.py), clamped at 0.")

    class _K(Kernel):
        is_stationary = False
        has_lengthscale = False

        def __init__(self, **kwargs):
            super().__init__(**kwargs)

        def forward(self, x1, x2, diag=False, **params):
            if diag:
                assert x1.size() == x2.size() and (x1 is x2 or torch.equal(x1, x2))
                return torch.ones(*x1.shape[:-2], x1.shape[-2], dtype=x1.dtype, device=x1.device)
            return self.covar_dist(x1, x2, **params)

        def covar_dist(self, x1, x2, last_dim_is_batch=False, **params):
            if last_dim_is_batch:
                x1 = x1.transpose(-1, -2).unsqueeze(-1)
                x2 = x2.transpose(-1, -2).unsqueeze(-1)
            return batch_sim(x1, x2)

    _K.__name__ = _K.__qualname__ = class_name
    _K.__doc__ = f"GPyTorch kernel wrapping ``batch_{name}_sim`` (drop-in for gauche's {class_name})."
    return batch_sim, _K


batch_dice_sim, DiceKernel = _make("dice", "DiceKernel")
batch_braun_blanquet_sim, BraunBlanquetKernel = _make("braun_blanquet", "BraunBlanquetKernel")
batch_faith_sim, FaithKernel = _make("faith", "FaithKernel")
batch_forbes_sim, ForbesKernel = _make("forbes", "ForbesKernel")
batch_inner_product_sim, InnerProductKernel = _make("inner_product", "InnerProductKernel")
batch_intersection_sim, IntersectionKernel = _make("intersection", "IntersectionKernel")
batch_otsuka_sim, OtsukaKernel = _make("otsuka", "OtsukaKernel")
batch_rand_sim, RandKernel = _make("rand", "RandKernel")
batch_rogers_tanimoto_sim, RogersTanimotoKernel = _make("rogers_tanimoto", "RogersTanimotoKernel")
batch_russell_rao_sim, RussellRaoKernel = _make("russell_rao", "RussellRaoKernel")
batch_sogenfrei_sim, SogenfreiKernel = _make("sogenfrei", "SogenfreiKernel")
batch_sokal_sneath_sim, SokalSneathKernel = _make("sokal_sneath", "SokalSneathKernel")

SIBLINGS = {
    "dice": (batch_dice_sim, DiceKernel), "braun_blanquet": (batch_braun_blanquet_sim, BraunBlanquetKernel),
    "faith": (batch_faith_sim, FaithKernel), "forbes": (batch_forbes_sim, ForbesKernel),
    "inner_product": (batch_inner_product_sim, InnerProductKernel),
    "intersection": (batch_intersection_sim, IntersectionKernel), "otsuka": (batch_otsuka_sim, OtsukaKernel),
    "rand": (batch_rand_sim, RandKernel), "rogers_tanimoto": (batch_rogers_tanimoto_sim, RogersTanimotoKernel),
    "russell_rao": (batch_russell_rao_sim, RussellRaoKernel), "sogenfrei": (batch_sogenfrei_sim, SogenfreiKernel),
    "sokal_sneath": (batch_sokal_sneath_sim, SokalSneathKernel),
}
