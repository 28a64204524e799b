"""Tensor-level entry points: shape normalisation + kernel-family dispatch over the C ABI.

``tanimoto_similarity`` / ``minmax_similarity`` are the arithmetic under
``batch_tanimoto_sim`` (gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:11-46) and
``batch_minmax_sim`` (gauche/kernels/fingerprint_kernels/minmax_kernel.py:10-44): same
argument meaning, same broadcasting of leading batch dimensions (``torch.matmul`` rules),
same error behaviour, output dtype/device of the inputs.

Kernel families (chosen from the packer's classification, never from a CPU fallback):
  binary  -> tcgen05 FP4 Gram: ``mxf4x1`` [default] ready-made e2m1 operands, ``bx`` packed bit words expanded to e2m1
             inside the SM (no expanded copy in HBM); tcgen05 int8 Gram (``umma2`` CTA pairs, ``umma2x1``) or ``__popc``
             CUDA-core (``popc``)
  count   -> the tcgen05 int8 kernels with u8 operands (``umma2`` [default]) or dp4a CUDA-core Gram (``u8``)
  real    -> dense fp32/fp64 CUDA-core Gram (tolerance-level parity)
MinMax always uses the u8 absolute-difference kernel for binary/count inputs.
"""
from __future__ import annotations

import itertools
import os
from typing import Optional, Sequence

import torch

from . import _lib
from .packed import PackedFingerprints, pack, pack_cache, stream_ptr

_FLOAT_OUT = {torch.float32: _lib.GK_F32, torch.float64: _lib.GK_F64}


def _engine_default() -> str:
    return os.environ.get("GAUCHE_B200_ENGINE", "auto")


def _require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.GaucheB200Error(
            "gauche_b200 needs a CUDA device (sm_100a); there is no CPU or PyTorch fallback"
        )
    return torch.device("cuda", torch.cuda.current_device())


def _check_inputs(x1: torch.Tensor, x2: torch.Tensor) -> None:
    # reference: tanimoto_kernel.py:33-34 / minmax_kernel.py:29-30
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    if x1.shape[-1] != x2.shape[-1]:
        raise RuntimeError(
            f"x1 and x2 must have the same feature dimension, got {x1.shape[-1]} and {x2.shape[-1]}"
        )
    if x1.dtype != x2.dtype:
        # torch.matmul / torch.cdist reject mixed dtypes
        raise RuntimeError(f"expected x1 and x2 to have the same dtype, got {x1.dtype} and {x2.dtype}")
    if x1.device != x2.device:
        raise RuntimeError(f"expected x1 and x2 on the same device, got {x1.device} and {x2.device}")


def _collapse_expanded(x: torch.Tensor) -> torch.Tensor:
    """Drop stride-0 (``expand``-ed) leading batch dims to size 1 so they are packed only once."""
    for d in range(x.ndim - 2):
        if x.shape[d] > 1 and x.stride(d) == 0:
            x = x.narrow(d, 0, 1)
    return x


def _needs_grad(x1: torch.Tensor, x2: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x1.requires_grad or x2.requires_grad)


class _TanimotoRealFn(torch.autograd.Function):
    """Differentiable Tanimoto on real-valued 2-D CUDA inputs (learnable inducing points, input warping;
    SURVEY.md §3.4-3.5).  Forward: ``gk_tanimoto_gram_real``.  Backward: ``gk_tanimoto_bwd_coeff`` recomputes
    the tile statistics and emits ``W = G*dK/ds`` plus the norm-gradient sums; the remaining two products
    ``W @ x2`` / ``W^T @ x1`` are plain library GEMMs."""

    @staticmethod
    def forward(ctx, x1, x2, eps):
        lib = _lib.load()
        x1c, x2c = x1.contiguous(), x2.contiguous()
        out = torch.empty((x1c.shape[0], x2c.shape[0]), dtype=x1c.dtype, device=x1c.device)
        if out.numel() > 0:
            with torch.cuda.device(x1c.device):
                _launch_real(lib, "tanimoto", x1c, x2c, out, eps, stream_ptr(x1c.device))
        ctx.save_for_backward(x1c, x2c)
        ctx.eps = eps
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable     # raw kernels: no double backward through W / ra / cb
    def backward(ctx, grad_out):
        x1, x2 = ctx.saved_tensors
        n, m, d = x1.shape[0], x2.shape[0], x1.shape[1]
        g = grad_out.contiguous()
        W = torch.empty_like(g)
        ra = torch.empty((n,), dtype=x1.dtype, device=x1.device)
        cb = torch.empty((m,), dtype=x1.dtype, device=x1.device)
        if n > 0 and m > 0:
            lib = _lib.load()
            with torch.cuda.device(x1.device):
                rc = lib.gk_tanimoto_bwd_coeff(x1.data_ptr(), max(d, 1), n, x2.data_ptr(), max(d, 1), m, d,
                                               g.data_ptr(), max(m, 1), W.data_ptr(), max(m, 1), ra.data_ptr(),
                                               cb.data_ptr(), _FLOAT_OUT[x1.dtype], ctx.eps, stream_ptr(x1.device))
            _lib.check(rc, "gk_tanimoto_bwd_coeff")
        else:
            ra.zero_(); cb.zero_(); W.zero_()
        g1 = g2 = None
        if ctx.needs_input_grad[0]:
            g1 = W @ x2 + 2.0 * x1 * ra[:, None]
        if ctx.needs_input_grad[1]:
            g2 = W.t() @ x1 + 2.0 * x2 * cb[:, None]
        return g1, g2, None


class _MinMaxRealFn(torch.autograd.Function):
    """Differentiable MinMax on real-valued 2-D CUDA inputs.  Forward: ``gk_minmax_gram_real``; backward:
    ``gk_minmax_bwd`` (coefficient kernel + two sign contractions, all CUDA)."""

    @staticmethod
    def forward(ctx, x1, x2, eps):
        lib = _lib.load()
        x1c, x2c = x1.contiguous(), x2.contiguous()
        out = torch.empty((x1c.shape[0], x2c.shape[0]), dtype=x1c.dtype, device=x1c.device)
        if out.numel() > 0:
            with torch.cuda.device(x1c.device):
                _launch_real(lib, "minmax", x1c, x2c, out, eps, stream_ptr(x1c.device))
        ctx.save_for_backward(x1c, x2c)
        ctx.eps = eps
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, grad_out):
        x1, x2 = ctx.saved_tensors
        n, m, d = x1.shape[0], x2.shape[0], x1.shape[1]
        g = grad_out.contiguous()
        g1 = torch.zeros_like(x1) if ctx.needs_input_grad[0] else None
        g2 = torch.zeros_like(x2) if ctx.needs_input_grad[1] else None
        if n > 0 and m > 0 and d > 0:
            cd = torch.empty_like(g)
            rs = torch.empty((n,), dtype=x1.dtype, device=x1.device)
            cs = torch.empty((m,), dtype=x1.dtype, device=x1.device)
            with torch.cuda.device(x1.device):
                rc = _lib.load().gk_minmax_bwd(
                    x1.data_ptr(), d, n, x2.data_ptr(), d, m, d, g.data_ptr(), m, cd.data_ptr(), m,
                    rs.data_ptr(), cs.data_ptr(), g1.data_ptr() if g1 is not None else None, d,
                    g2.data_ptr() if g2 is not None else None, d, _FLOAT_OUT[x1.dtype], ctx.eps,
                    stream_ptr(x1.device))
            _lib.check(rc, "gk_minmax_bwd")
        return g1, g2, None


def _with_grad(fn, x1: torch.Tensor, x2: torch.Tensor, eps: float) -> torch.Tensor:
    if not (x1.is_cuda and x1.dtype in _FLOAT_OUT):
        raise NotImplementedError("gauche_b200: gradients need float32/float64 CUDA inputs")
    if x1.ndim == 2 and x2.ndim == 2:
        return fn.apply(x1, x2, eps)
    batch = torch.broadcast_shapes(x1.shape[:-2], x2.shape[:-2])
    a = x1.expand(*batch, *x1.shape[-2:]).reshape(-1, *x1.shape[-2:])
    b = x2.expand(*batch, *x2.shape[-2:]).reshape(-1, *x2.shape[-2:])
    out = torch.stack([fn.apply(a[i], b[i], eps) for i in range(a.shape[0])])
    return out.reshape(*batch, x1.shape[-2], x2.shape[-2])


class _Operand:
    """A batched operand ``[..., rows, D]`` flattened to one packed ``[B*rows, D]`` matrix."""

    def __init__(self, x: torch.Tensor, use_cache: bool):
        xc = _collapse_expanded(x)
        self.batch_shape = tuple(xc.shape[:-2])
        self.rows, self.dim = xc.shape[-2], xc.shape[-1]
        self.nbatch = 1
        for s in self.batch_shape:
            self.nbatch *= s
        flat = xc if xc.ndim == 2 else xc.reshape(self.nbatch * self.rows, self.dim)
        self.flat = flat
        # identity-keyed cache only makes sense for the caller's own 2-D tensor object
        self.packed: PackedFingerprints = (
            pack_cache.get(flat) if (use_cache and flat is x) else pack(flat)
        )

    def batch_index(self, out_index: Sequence[int], out_ndim: int) -> int:
        """Linear batch index of this operand for a broadcast output batch index."""
        lin = 0
        nb = len(self.batch_shape)
        for d, size in enumerate(self.batch_shape):
            i = out_index[out_ndim - nb + d] if size > 1 else 0
            lin = lin * size + i
        return lin

    def entry(self, b: int) -> PackedFingerprints:
        if self.nbatch == 1:
            return self.packed
        return self.packed.rows_slice(b * self.rows, (b + 1) * self.rows)

    def dense_entry(self, b: int) -> torch.Tensor:
        return self.flat[b * self.rows:(b + 1) * self.rows]


TANIMOTO_ENGINES = ("bxh", "mxf4x1", "bx", "umma2", "umma2x1", "popc", "u8")
last_launch = {"tanimoto": None, "minmax": None}    # C entry point of the most recent Tanimoto launch (smoke / tests assert on it)


def _launch_tanimoto(lib, engine: str, a: PackedFingerprints, b: PackedFingerprints, out: torch.Tensor,
                     out_code: int, eps: float, stream: int) -> None:
    """One Gram launch; ``out`` may have any row pitch (the tensor-core epilogue has an LSU path for pitches the
    TMA cannot address)."""
    n, m = a.rows, b.rows
    ldo = out.stride(0) if n > 1 else max(m, 1)
    if engine == "bxh" and out_code == _lib.GK_F32 and 1e-12 <= eps <= 1.0:
        # hybrid: A = permuted bit words expanded into tensor memory, B = ready-made e2m1 by TMA (fp32 Gram; the other
        # outputs are bound by their epilogues, not by the operand pipeline, and take the e2m1 engine)
        name = "gk_tanimoto_gram_hybrid"
        ap, b4 = a.bitsp(), b.q4()
        rc = lib.gk_tanimoto_gram_hybrid(ap.data_ptr(), ap.stride(0) if n > 1 else a.k4_bytes // 16, a.rowsq.data_ptr(), n,
                                         b4.data_ptr(), b.k4_bytes, b.rowsq.data_ptr(), m,
                                         a.k4_bytes // 16, b.k4_bytes, out.data_ptr(), ldo, out_code, eps, stream)
    elif engine == "bx":   # packed bits all the way into the SM, expanded to e2m1 by the kernel's expander warps
        name = "gk_tanimoto_gram_bits"
        rc = lib.gk_tanimoto_gram_bits(a.bits.data_ptr(), a.bits.stride(0) if n > 1 else a.kp // 32, a.rowsq.data_ptr(), n,
                                       b.bits.data_ptr(), b.bits.stride(0) if m > 1 else b.kp // 32, b.rowsq.data_ptr(), m,
                                       a.kp // 32, out.data_ptr(), ldo, out_code, eps, stream)
    elif engine in ("umma2", "umma2x1"):
        name = "gk_tanimoto_gram_umma2"
        rc = lib.gk_tanimoto_gram_umma2(a.q8.data_ptr(), a.kp, a.rowsq.data_ptr(), n,
                                        b.q8.data_ptr(), b.kp, b.rowsq.data_ptr(), m,
                                        a.kp, out.data_ptr(), ldo, out_code, eps,
                                        2 if engine == "umma2" else 1, stream)
    elif engine in ("mxf4x1", "bxh"):    # ready-made e2m1 operands in HBM
        name = "gk_tanimoto_gram_mxf4"
        a4, b4 = a.q4(), (a.q4() if b is a else b.q4())
        rc = lib.gk_tanimoto_gram_mxf4(a4.data_ptr(), a.k4_bytes, a.rowsq.data_ptr(), n,
                                       b4.data_ptr(), b.k4_bytes, b.rowsq.data_ptr(), m,
                                       a.k4_bytes, out.data_ptr(), ldo, out_code, eps, stream)
    elif engine == "popc":
        name = "gk_tanimoto_gram_popc"
        rc = lib.gk_tanimoto_gram_popc(a.bits.data_ptr(), a.kp // 32, a.rowsq.data_ptr(), n,
                                       b.bits.data_ptr(), b.kp // 32, b.rowsq.data_ptr(), m,
                                       a.kp // 32, out.data_ptr(), ldo, out_code, eps, stream)
    elif engine == "u8":
        name = "gk_tanimoto_gram_u8"
        rc = lib.gk_tanimoto_gram_u8(a.q8.data_ptr(), a.kp, a.rowsq.data_ptr(), n,
                                     b.q8.data_ptr(), b.kp, b.rowsq.data_ptr(), m,
                                     a.kp, out.data_ptr(), ldo, out_code, eps, stream)
    else:
        raise ValueError(f"unknown Tanimoto engine {engine!r}")
    _lib.check(rc, name)
    last_launch["tanimoto"] = name


MINMAX_TC_MAX_PLANES = 48   # beyond this the VABSDIFF4 CUDA-core kernel wins (cost grows with the depth)
MINMAX_SPARSE_MIN_PLANES = 3    # from here on empty thermometer blocks are worth skipping
MINMAX_PAIRS = True             # block-sparse MinMax on CTA pairs (measured: profiles/r2_minmax_pairs.log)
MINMAX_PAIR_MIN_ROWS = 4096
MINMAX_SYM_PAIR_MIN_ROWS = 8192
MINMAX_SYMMETRIC = True         # x1 is x2: upper-triangle tiles + mirrored stores (exact: the MinMax Gram is bitwise symmetric)


def _launch_minmax(lib, a: PackedFingerprints, b: PackedFingerprints, out: torch.Tensor, out_code: int,
                   eps: float, stream: int, engine: Optional[str] = None) -> None:
    n, m = a.rows, b.rows
    ldo = out.stride(0) if n > 1 else max(m, 1)
    planes = max(a.maxval, b.maxval, 1)
    use_tc = engine != "u8" and planes <= MINMAX_TC_MAX_PLANES and planes * a.k4_bytes * 2 < (1 << 24)
    if engine in ("tc", "tc-dense", "tc-pair", "tc-single", "tc-full") and not use_tc:
        raise ValueError(f"MinMax tensor-core path unavailable (planes={planes})")
    if use_tc:
        # thermometer planes on the FP4 tensor cores: sum_c min(a_c,b_c) = sum_t <[a>=t],[b>=t]>
        a4 = a.q4t(planes)
        b4 = a4 if b is a else b.q4t(planes)
        kb = planes * a.k4_bytes
        # high thermometer planes are mostly empty blocks: skip them (exact); needs the packed-fp epilogue's eps range
        if planes >= MINMAX_SPARSE_MIN_PLANES and 1e-12 <= eps <= 1.0 and engine != "tc-dense":
            # CTA pairs (256-row tiles, half of the B tile staged per CTA) for problems with enough row tiles: the MinMax
            # kernel is bound by operand ingress (L2 -> SM), which pairs cut by a third
            sym = b is a and MINMAX_SYMMETRIC and engine != "tc-full" and out_code in (_lib.GK_F32, _lib.GK_I32)
            # the symmetric kernel has half the tiles: CTA pairs pay from 8k rows on (profiles/r2_minmax_symmetric.log)
            pair_rows = MINMAX_SYM_PAIR_MIN_ROWS if sym else MINMAX_PAIR_MIN_ROWS
            cg = 2 if (engine == "tc-pair" or (engine != "tc-single" and MINMAX_PAIRS and n >= pair_rows)) else 1
            occ_a = a.occupancy(planes, 128 * cg)
            occ_b = b.occupancy(planes, 224)
            if sym:
                # K(x, x): the reference's MinMax Gram is bitwise symmetric -> compute the upper triangle of tiles only and
                # store every chunk mirrored as well (half the tensor work, same bytes written)
                rc = lib.gk_minmax_gram_mxf4_sym(a4.data_ptr(), a4.stride(0) if n > 1 else kb, a.rowsum.data_ptr(), n, kb,
                                                 occ_a.data_ptr(), occ_b.data_ptr(), out.data_ptr(), ldo, out_code, eps, cg,
                                                 stream)
                _lib.check(rc, "gk_minmax_gram_mxf4_sym")
                last_launch["minmax"] = "gk_minmax_gram_mxf4_sym"
                return
            last_launch["minmax"] = "gk_minmax_gram_mxf4_sparse"
            rc = lib.gk_minmax_gram_mxf4_sparse(a4.data_ptr(), a4.stride(0) if n > 1 else kb, a.rowsum.data_ptr(), n,
                                                b4.data_ptr(), b4.stride(0) if m > 1 else kb, b.rowsum.data_ptr(), m,
                                                kb, occ_a.data_ptr(), occ_b.data_ptr(), out.data_ptr(), ldo, out_code,
                                                eps, cg, stream)
            _lib.check(rc, "gk_minmax_gram_mxf4_sparse")
            return
        rc = lib.gk_minmax_gram_mxf4(a4.data_ptr(), a4.stride(0) if n > 1 else kb, a.rowsum.data_ptr(), n,
                                     b4.data_ptr(), b4.stride(0) if m > 1 else kb, b.rowsum.data_ptr(), m,
                                     kb, out.data_ptr(), ldo, out_code, eps, stream)
        _lib.check(rc, "gk_minmax_gram_mxf4")
        return
    rc = lib.gk_minmax_gram_u8(a.q8.data_ptr(), a.kp, a.rowsum.data_ptr(), n,
                               b.q8.data_ptr(), b.kp, b.rowsum.data_ptr(), m,
                               a.kp, out.data_ptr(), ldo, out_code, eps, stream)
    _lib.check(rc, "gk_minmax_gram_u8")


def _launch_real(lib, metric: str, x1: torch.Tensor, x2: torch.Tensor, out: torch.Tensor, eps: float,
                 stream: int) -> None:
    n, m, d = x1.shape[0], x2.shape[0], x1.shape[1]
    fn = lib.gk_tanimoto_gram_real if metric == "tanimoto" else lib.gk_minmax_gram_real
    rc = fn(x1.data_ptr(), x1.stride(0) if n > 1 else max(d, 1), n,
            x2.data_ptr(), x2.stride(0) if m > 1 else max(d, 1), m, d,
            out.data_ptr(), out.stride(0) if n > 1 else max(m, 1), _FLOAT_OUT[out.dtype], eps, stream)
    _lib.check(rc, f"gk_{metric}_gram_real")


def resolve_engine(metric: str, kind: str, engine: Optional[str]) -> str:
    engine = engine or _engine_default()
    if kind == "real":
        return "real"
    if metric == "minmax":
        return "u8"      # kernel family; tensor-core (thermometer) vs VABSDIFF4 is chosen per launch
    if engine == "auto":
        # bits: the hybrid kind::mxf4 kernel (A = bit words expanded into tensor memory, B = ready-made e2m1; exact; measured
        # fastest at every size, profiles/r2_bxh_hybrid_kernel.log); counts: u8 operands on kind::i8 CTA pairs
        return "bxh" if kind == "binary" else "umma2"
    if engine not in TANIMOTO_ENGINES:
        raise ValueError(f"unknown engine {engine!r} (expected auto|{'|'.join(TANIMOTO_ENGINES)})")
    if kind != "binary":
        # bit operands (bxh / bx / mxf4x1 / popc) hold bits only; counts stay on the integer kernels
        return "u8" if engine in ("popc", "u8") else ("umma2x1" if engine == "umma2x1" else "umma2")
    return engine


def _broadcast(a: tuple, b: tuple) -> tuple:
    """``torch.broadcast_shapes`` with the common cases (no batch dims, equal shapes) kept off its slow Python path:
    small Grams are launch-bound and it was a third of the host time of a cached call."""
    if a == b or not b:
        return a
    if not a:
        return b
    return tuple(torch.broadcast_shapes(a, b))


def _similarity(metric: str, x1: torch.Tensor, x2: torch.Tensor, eps: float, engine: Optional[str],
                raw_counts: bool = False) -> torch.Tensor:
    _check_inputs(x1, x2)
    if _needs_grad(x1, x2):
        if raw_counts:
            raise NotImplementedError("gauche_b200: raw integer counts are not differentiable")
        if metric == "minmax" and not x1.dtype.is_floating_point:
            raise RuntimeError(f"cdist only supports floating-point dtypes, X1 got: {x1.dtype}")
        return _with_grad(_TanimotoRealFn if metric == "tanimoto" else _MinMaxRealFn, x1, x2, eps)
    if metric == "minmax" and not x1.dtype.is_floating_point:
        # reference: torch.cdist (minmax_kernel.py:36) rejects integer tensors
        raise RuntimeError(f"cdist only supports floating-point dtypes, X1 got: {x1.dtype}")

    in_device = x1.device
    if not x1.is_cuda:
        dev = _require_cuda()
        same = x2 is x1
        x1 = x1.to(dev, non_blocking=True)
        x2 = x1 if same else x2.to(dev, non_blocking=True)
    dev = x1.device
    lib = _lib.load()

    # output dtype: input float dtype; integer inputs promote to the default float32 like
    # `int_tensor + 1e-6` does in the reference (tanimoto_kernel.py:40)
    if x1.dtype in (torch.float32, torch.float64):
        out_dtype = x1.dtype
        final_dtype = x1.dtype
    elif x1.dtype.is_floating_point:
        out_dtype, final_dtype = torch.float32, x1.dtype
    else:
        out_dtype = final_dtype = torch.float32
    if raw_counts:
        out_dtype = final_dtype = torch.int32

    use_cache = pack_cache.enabled
    op1 = _Operand(x1, use_cache)
    op2 = op1 if x2 is x1 else _Operand(x2, use_cache)
    batch = _broadcast(tuple(op1.batch_shape), tuple(op2.batch_shape))
    full_batch = _broadcast(tuple(x1.shape[:-2]), tuple(x2.shape[:-2]))
    n, m = op1.rows, op2.rows
    out = torch.empty(tuple(batch) + (n, m), dtype=out_dtype, device=dev)

    if out.numel() > 0 and op1.dim > 0:
        kinds = {op1.packed.kind, op2.packed.kind}
        kind = "real" if "real" in kinds else ("count" if "count" in kinds else "binary")
        if raw_counts and kind == "real":
            raise ValueError("raw integer counts are only defined for binary / count fingerprints")
        eng = resolve_engine(metric, kind, engine)
        out_code = _lib.GK_I32 if raw_counts else _FLOAT_OUT[out_dtype]
        stream = stream_ptr(dev)
        if eng == "real":
            d1 = op1.flat if op1.flat.dtype == out_dtype else op1.flat.to(out_dtype)
            d2 = d1 if op2 is op1 else (op2.flat if op2.flat.dtype == out_dtype else op2.flat.to(out_dtype))
            if d1.stride(-1) != 1:
                d1 = d1.contiguous()
            if d2.stride(-1) != 1:
                d2 = d2.contiguous()
        out_flat = out.reshape((-1, n, m))
        with torch.cuda.device(dev):
            for lin, idx in enumerate(itertools.product(*[range(s) for s in batch])):
                b1 = op1.batch_index(idx, len(batch))
                b2 = op2.batch_index(idx, len(batch))
                o = out_flat[lin]
                if eng == "real":
                    _launch_real(lib, metric, d1[b1 * n:(b1 + 1) * n], d2[b2 * m:(b2 + 1) * m], o, eps, stream)
                elif metric == "tanimoto":
                    _launch_tanimoto(lib, eng, op1.entry(b1), op2.entry(b2), o, out_code, eps, stream)
                else:
                    _launch_minmax(lib, op1.entry(b1), op2.entry(b2), o, out_code, eps, stream,
                                   engine if engine in ("tc", "tc-dense", "tc-pair", "tc-single", "tc-full", "u8") else None)
    elif out.numel() > 0:
        # D == 0: dot = norms = 0 -> eps/eps = 1 exactly as the reference computes it
        out.fill_(0 if raw_counts else 1)

    if tuple(batch) != tuple(full_batch):
        out = out.expand(tuple(full_batch) + (n, m))
    if final_dtype != out.dtype:
        out = out.to(final_dtype)
    if out.device != in_device:
        out = out.to(in_device)
    return out


def tanimoto_similarity(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6,
                        engine: Optional[str] = None) -> torch.Tensor:
    return _similarity("tanimoto", x1, x2, eps, engine)


def minmax_similarity(x1: torch.Tensor, x2: torch.Tensor, eps: float = 1e-6,
                      engine: Optional[str] = None) -> torch.Tensor:
    return _similarity("minmax", x1, x2, eps, engine)


def intersection_counts(x1: torch.Tensor, x2: torch.Tensor, engine: Optional[str] = None) -> torch.Tensor:
    """Exact int32 ``<x1_i, x2_j>`` (``|A∩B|`` for bits) straight from the integer accumulators."""
    return _similarity("tanimoto", x1, x2, 0.0, engine, raw_counts=True)


def l1_distances(x1: torch.Tensor, x2: torch.Tensor, engine: Optional[str] = None) -> torch.Tensor:
    """Exact int32 ``sum_k |x1_ik - x2_jk|`` of the MinMax kernel (``engine``: ``tc`` thermometer planes
    on the FP4 tensor cores [``tc-dense``: without the block-sparse k loop], ``u8`` VABSDIFF4 CUDA cores,
    default: chosen from the largest count)."""
    return _similarity("minmax", x1, x2, 0.0, engine, raw_counts=True)
