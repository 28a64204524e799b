"""Exact-integer numpy restatement of the reference arithmetic (test infrastructure only).

Follows, line by line,
  gauche/kernels/fingerprint_kernels/tanimoto_kernel.py:33-46   (batch_tanimoto_sim)
  gauche/kernels/fingerprint_kernels/minmax_kernel.py:29-44     (batch_minmax_sim)
but computes the contraction in exact integers first (intersection counts / L1 distances) and
then replays the reference's floating-point operations one at a time, in the reference's order
and dtype.  For integer-valued fingerprints whose counts are below 2**24 (fp32) / 2**53 (fp64)
the reference's float matmul / sum / cdist are themselves exact, so the two agree bit for bit.
"""
from __future__ import annotations

import numpy as np


def _as_int_matrix(x: np.ndarray) -> np.ndarray:
    xi = np.asarray(x)
    if not np.all(xi == np.floor(xi)):
        raise ValueError("exact oracle expects integer-valued fingerprints")
    return xi.astype(np.float64)  # float64 BLAS matmul is exact for these magnitudes


def tanimoto_counts(x1: np.ndarray, x2: np.ndarray):
    """Exact ``inter[i,j] = <x1_i, x2_j>``, ``n1[i] = sum x1_i**2``, ``n2[j]`` as int64.
    Mirrors tanimoto_kernel.py:36-38 (matmul, sum of squares)."""
    a, b = _as_int_matrix(x1), _as_int_matrix(x2)
    inter = np.rint(a @ np.swapaxes(b, -1, -2)).astype(np.int64)
    n1 = np.rint((a * a).sum(-1)).astype(np.int64)
    n2 = np.rint((b * b).sum(-1)).astype(np.int64)
    return inter, n1, n2


def tanimoto_sim_from_counts(inter, n1, n2, dtype=np.float32, eps: float = 1e-6) -> np.ndarray:
    """Replays tanimoto_kernel.py:40-46 in ``dtype``: (dot+eps) / (((eps+n1)+n2ᵀ)-dot), clamp_min 0."""
    dt = np.dtype(dtype).type
    e = dt(eps)
    dot = inter.astype(dt)
    r = (e + n1.astype(dt))[..., :, None]          # eps + x1_norm
    c = n2.astype(dt)[..., None, :]                # transpose(x2_norm)
    num = dot + e
    den = (r + c) - dot
    with np.errstate(divide="ignore", invalid="ignore"):
        q = num / den
    return np.where(q < 0, dt(0), q).astype(dt)


def tanimoto_sim(x1, x2, eps: float = 1e-6, dtype=None) -> np.ndarray:
    """batch_tanimoto_sim for integer-valued inputs; output dtype = input float dtype, float32 for
    integer inputs (type promotion of ``int_tensor + eps``)."""
    x1, x2 = np.asarray(x1), np.asarray(x2)
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    if dtype is None:
        dtype = x1.dtype if x1.dtype.kind == "f" else np.float32
    inter, n1, n2 = tanimoto_counts(x1, x2)
    return tanimoto_sim_from_counts(inter, n1, n2, dtype, eps)


def minmax_l1(x1: np.ndarray, x2: np.ndarray, chunk: int = 64):
    """Exact ``d[i,j] = sum_k |x1_ik - x2_jk|`` (== torch.cdist(p=1), minmax_kernel.py:36) and the
    L1 norms ``s1[i] = sum_k x1_ik`` (:33-34), all int64.  2-D inputs."""
    a = np.asarray(x1)
    b = np.asarray(x2)
    if not (np.all(a == np.floor(a)) and np.all(b == np.floor(b))):
        raise ValueError("exact oracle expects integer-valued fingerprints")
    a = a.astype(np.int64)
    b = b.astype(np.int64)
    n, m = a.shape[0], b.shape[0]
    d = np.empty((n, m), dtype=np.int64)
    for i0 in range(0, n, chunk):
        blk = a[i0:i0 + chunk, None, :]
        d[i0:i0 + chunk] = np.abs(blk - b[None, :, :]).sum(-1)
    return d, a.sum(-1), b.sum(-1)


def minmax_sim_from_counts(d, s1, s2, dtype=np.float64, eps: float = 1e-6) -> np.ndarray:
    """Replays minmax_kernel.py:35-44: s = s1 + s2ᵀ; ((s-d)+eps) / ((s+d)+eps), clamp_min 0."""
    dt = np.dtype(dtype).type
    e = dt(eps)
    s = s1.astype(dt)[:, None] + s2.astype(dt)[None, :]
    dist = d.astype(dt)
    num = (s - dist) + e
    den = (s + dist) + e
    with np.errstate(divide="ignore", invalid="ignore"):
        q = num / den
    return np.where(q < 0, dt(0), q).astype(dt)


def minmax_sim(x1, x2, eps: float = 1e-6) -> np.ndarray:
    """batch_minmax_sim for integer-valued float inputs (2-D or with equal leading batch dims)."""
    x1, x2 = np.asarray(x1), np.asarray(x2)
    if x1.ndim < 2 or x2.ndim < 2:
        raise ValueError("Tensors must have a batch dimension")
    if x1.dtype.kind != "f":
        raise RuntimeError("cdist only supports floating-point dtypes")
    if x1.ndim == 2 and x2.ndim == 2:
        d, s1, s2 = minmax_l1(x1, x2)
        return minmax_sim_from_counts(d, s1, s2, x1.dtype, eps)
    batch = np.broadcast_shapes(x1.shape[:-2], x2.shape[:-2])
    a = np.broadcast_to(x1, batch + x1.shape[-2:])
    b = np.broadcast_to(x2, batch + x2.shape[-2:])
    out = np.empty(batch + (x1.shape[-2], x2.shape[-2]), dtype=x1.dtype)
    for idx in np.ndindex(*batch):
        d, s1, s2 = minmax_l1(a[idx], b[idx])
        out[idx] = minmax_sim_from_counts(d, s1, s2, x1.dtype, eps)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Operand layouts of the hybrid Gram kernel (test infrastructure for gk_permute_bits_kblock / gk_expand_bits_e2m1; the
# layouts are this repo's own — the reference has dense tensors only, tanimoto_kernel.py:36).
def permute_bits_kblock(bits: np.ndarray) -> np.ndarray:
    """``bits`` uint32 ``[rows, words]`` (bit i of word w = fingerprint bit 32w + i) -> uint32 ``[rows, ceil(words/8)*8]``:
    inside every 256-bit k-block, bit 8j+i of word 2s+h moves to bit 4i+s of word 4h+j (h < 2, j < 4, s < 4, i < 8)."""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    rows, words = bits.shape
    wp = (words + 7) // 8 * 8
    src = np.zeros((rows, wp), dtype=np.uint32)
    src[:, :words] = bits
    out = np.zeros_like(src)
    for kb in range(wp // 8):
        for h in range(2):
            for j in range(4):
                o = np.zeros(rows, dtype=np.uint32)
                for s in range(4):
                    w = src[:, kb * 8 + 2 * s + h]
                    for i in range(8):
                        o |= ((w >> np.uint32(8 * j + i)) & np.uint32(1)) << np.uint32(4 * i + s)
                out[:, kb * 8 + 4 * h + j] = o
    return out


def expand_kblock_rows(words8: np.ndarray) -> np.ndarray:
    """The kernel's five-op expansion of one k-block row (eight bit words) into 32 words of e2m1 nibbles: chunk 2s+h, word j
    = (word[4h+j] >> sh_s) & mask_s with masks 0x1111.., 0x2222.., 0x4444.., 0x4444.. and sh = 0, 0, 0, 1.  Returns the
    NON-ZERO pattern per nibble (uint8 [.., 256]) in the order the MMA sees the K elements."""
    words8 = np.asarray(words8, dtype=np.uint32)
    masks = (0x11111111, 0x22222222, 0x44444444, 0x44444444)
    shifts = (0, 0, 0, 1)
    out = np.zeros(words8.shape[:-1] + (256,), dtype=np.uint8)
    for s in range(4):
        for h in range(2):
            for j in range(4):
                v = (words8[..., 4 * h + j] >> np.uint32(shifts[s])) & np.uint32(masks[s])
                for i in range(8):
                    out[..., 64 * s + 32 * h + 8 * j + i] = ((v >> np.uint32(4 * i)) & np.uint32(0xF)) != 0
    return out
