"""The other twelve fingerprint kernels of ``gauche.kernels.fingerprint_kernels`` on the B200 path.

Bit / count fingerprints: each kernel is an EPILOGUE of the tensor-core Gram kernel (``gk_sibling_gram``): the exact
integer dot products stay in TMEM and the similarity tile is written once, evaluated in the reference's operation order,
so the results are bit-identical.  Real-valued inputs take the dense CUDA-core path (``gk_sibling_gram_real``,
tolerance-level parity).  Leading batch dimensions broadcast like ``torch.matmul`` does in the reference, and
``last_dim_is_batch`` works through them.

The reference's shape quirks are reproduced, not fixed — and they are data, not code: ``d`` (common zeros of the LAST rows /
last batch entry), the Braun-Blanquet denominator ``max(x1_norm[-1], x2_norm[-1])`` and the untransposed ``x2_norm`` of
Rogers-Tanimoto / Sokal-Sneath are evaluated with the reference's own indexing expressions on the small norm tensors and
handed to the kernel as one integer per output row.  Reference: gauche/kernels/fingerprint_kernels/<name>_kernel.py.
"""
from __future__ import annotations

import math

import torch

from ... import _lib
from ..._kernel_base import Kernel
from ...ops import _check_inputs, _needs_grad, _require_cuda
from ...packed import pack, pack_cache, stream_ptr

_IDS = {name: i for i, name in enumerate([
    "dice", "braun_blanquet", "faith", "forbes", "inner_product", "intersection", "otsuka", "rand",
    "rogers_tanimoto", "russell_rao", "sogenfrei", "sokal_sneath"])}
# kernels whose reference ends with `.to(dtype=torch.double)`
_TO_DOUBLE = {"braun_blanquet", "forbes", "intersection", "otsuka", "rand", "rogers_tanimoto", "russell_rao",
              "sogenfrei", "sokal_sneath"}
_NEEDS_D = {"faith", "intersection", "rand", "rogers_tanimoto"}
_UNTRANSPOSED = {"rogers_tanimoto", "sokal_sneath"}
_CODES = {torch.float32: _lib.GK_F32, torch.float64: _lib.GK_F64}
last_launch = {"sibling": None}       # C entry point of the most recent launch (tests assert on it)


def _unbroadcast(t: torch.Tensor, batch: tuple, full_batch: tuple) -> torch.Tensor:
    """``t`` is ``full_batch + tail``; return the ``batch + tail`` tensor it was broadcast from (index 0 along every
    dimension that broadcasting added or stretched)."""
    idx = []
    lead = len(full_batch) - len(batch)
    for i in range(len(full_batch)):
        if i < lead:
            idx.append(0)
        elif batch[i - lead] == 1 and full_batch[i] > 1:
            idx.append(slice(0, 1))
        else:
            idx.append(slice(None))
    return t[tuple(idx)]


def _sibling_sim(name: str, x1: torch.Tensor, x2: torch.Tensor, eps: float) -> torch.Tensor:
    _check_inputs(x1, x2)
    if _needs_grad(x1, x2):
        raise NotImplementedError(f"gauche_b200: {name} kernel is not differentiable w.r.t. its inputs")
    if name == "intersection" and (bool(torch.any(x1 > 1)) or bool(torch.any(x2 > 1))):
        raise ValueError("Tensors must be binary-valued")        # intersection_kernel.py:36-37
    in_device = x1.device
    same = x2 is x1
    if not x1.is_cuda:
        dev = _require_cuda()
        x1 = x1.to(dev)
        x2 = x1 if same else x2.to(dev)
    dev = x1.device
    n, m, dim = x1.shape[-2], x2.shape[-2], x1.shape[-1]
    bs1, bs2 = tuple(x1.shape[:-2]), tuple(x2.shape[:-2])
    full_batch = tuple(torch.broadcast_shapes(bs1, bs2))
    nb = math.prod(full_batch)
    compute = x1.dtype if x1.dtype in _CODES else torch.float32      # ints promote like `int_tensor + eps`
    out_dtype = torch.float64 if name in _TO_DOUBLE else compute
    final_dtype = out_dtype if (name in _TO_DOUBLE or x1.dtype in _CODES or not x1.dtype.is_floating_point) else x1.dtype
    out = torch.empty((nb, n, m), dtype=out_dtype, device=dev)

    def finish(o):
        o = o.view(full_batch + (n, m))
        if o.dtype != final_dtype:
            o = o.to(final_dtype)
        return o if o.device == in_device else o.to(in_device)

    if out.numel() == 0:
        return finish(out)
    xa = x1 if dim > 0 else x1.new_zeros(bs1 + (n, 1))               # D == 0: all dot products and norms are zero
    xb = (xa if same else (x2 if dim > 0 else x2.new_zeros(bs2 + (m, 1))))
    a = xa.expand(full_batch + tuple(xa.shape[-2:])).reshape(nb * n, xa.shape[-1])
    b = a if (same and bs1 == full_batch) else xb.expand(full_batch + tuple(xb.shape[-2:])).reshape(nb * m, xb.shape[-1])
    cached = pack_cache.enabled
    p1 = pack_cache.get(a) if (cached and a is x1) else pack(a)
    p2 = p1 if b is a else (pack_cache.get(b) if (cached and b is x2) else pack(b))
    real = "real" in (p1.kind, p2.kind)
    stream = stream_ptr(dev)
    lib = _lib.load()
    if real:
        a = a if a.dtype == compute else a.to(compute)
        b = a if p2 is p1 else (b if b.dtype == compute else b.to(compute))
        a, b = a.contiguous(), b.contiguous()
        l1a = torch.empty((nb * n,), dtype=compute, device=dev)
        l1b = l1a if b is a else torch.empty((nb * m,), dtype=compute, device=dev)
        with torch.cuda.device(dev):
            for x, o in ((a, l1a),) + (() if b is a else ((b, l1b),)):
                _lib.check(lib.gk_rowsum_real(x.data_ptr(), x.stride(0) if x.shape[0] > 1 else max(x.shape[1], 1), x.shape[0],
                                              x.shape[1], _CODES[compute], o.data_ptr(), stream), "gk_rowsum_real")
        n1_full, n2_full = l1a.view(full_batch + (n, 1)), l1b.view(full_batch + (m, 1))
    else:
        n1_full, n2_full = p1.rowsum.view(full_batch + (n, 1)), p2.rowsum.view(full_batch + (m, 1))
    # the reference's norm tensors have the ORIGINAL batch shapes; its quirks index those
    x1_norm, x2_norm = _unbroadcast(n1_full, bs1, full_batch), _unbroadcast(n2_full, bs2, full_batch)
    rows_shape = full_batch + (n, 1)
    term_dtype = compute if real else torch.int32
    row_a = n1_full
    row_b = None
    if name in _UNTRANSPOSED:
        s = x1_norm + x2_norm                                          # `2 * x1_norm + 2 * x2_norm`: x2_norm NOT transposed
        if tuple(torch.broadcast_shapes(tuple(s.shape), rows_shape)) != rows_shape:
            raise RuntimeError(f"The size of tensor a ({n}) must match the size of tensor b ({m}) at non-singleton "
                               "dimension 0" if n != 1 else "gauche_b200: N == 1 < M is not supported for this kernel")
        row_a = torch.broadcast_to(s, rows_shape)
    if name in _NEEDS_D:
        d = torch.sum((x1[-1] == 0) & (x2[-1] == 0), dim=-1, keepdim=True)   # the reference's own expression (last rows only)
        row_b = torch.broadcast_to(d, rows_shape)
    elif name == "braun_blanquet":
        row_b = torch.broadcast_to(torch.max(x1_norm[-1], x2_norm[-1]), rows_shape)   # braun_blanquet_kernel.py:39
    row_a = row_a.to(term_dtype).reshape(nb, n).contiguous()
    row_b = None if row_b is None else row_b.to(term_dtype).reshape(nb, n).contiguous()
    col = n2_full.to(term_dtype).reshape(nb, m).contiguous()

    which = _IDS[name]
    with torch.cuda.device(dev):
        for i in range(nb):
            o = out[i]
            rb_ptr = None if row_b is None else row_b[i].data_ptr()
            if real:
                ai, bi = a[i * n:(i + 1) * n], b[i * m:(i + 1) * m]
                rc = lib.gk_sibling_gram_real(which, ai.data_ptr(), ai.stride(0) if n > 1 else max(ai.shape[1], 1),
                                              row_a[i].data_ptr(), rb_ptr, n, bi.data_ptr(),
                                              bi.stride(0) if m > 1 else max(bi.shape[1], 1), col[i].data_ptr(), m,
                                              ai.shape[1], dim, _CODES[compute], o.data_ptr(), max(m, 1), _CODES[out_dtype],
                                              eps, stream)
                entry = "gk_sibling_gram_real"
            else:
                e1 = p1 if nb == 1 else p1.rows_slice(i * n, (i + 1) * n)
                e2 = e1 if (p2 is p1 and n == m) else (p2 if nb == 1 else p2.rows_slice(i * m, (i + 1) * m))
                if p1.kind == "binary" and p2.kind == "binary":
                    kind, qa, qb, kb = 1, e1.q4(), e2.q4(), e1.k4_bytes
                else:
                    kind, qa, qb, kb = 0, e1.q8, e2.q8, e1.kp
                rc = lib.gk_sibling_gram(which, kind, qa.data_ptr(), qa.stride(0) if n > 1 else kb, row_a[i].data_ptr(), rb_ptr, n,
                                         qb.data_ptr(), qb.stride(0) if m > 1 else kb, col[i].data_ptr(), m, kb, dim,
                                         _CODES[compute], o.data_ptr(), max(m, 1), _CODES[out_dtype], eps, stream)
                entry = "gk_sibling_gram"
            _lib.check(rc, entry)
    last_launch["sibling"] = entry
    return finish(out)


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
