"""Packed fingerprint operands: the HBM data layout of the integer Gram kernels.

A dense ``[rows, D]`` fingerprint matrix (fp64 = 16 KB per 2048-bit molecule) is turned
once, by ``gk_pack_classify``, into

* ``bits``   ``[rows, Kp/32]`` int32 words, bit ``i`` of word ``w`` = ``x[., 32w+i] != 0``  (256 B / molecule)
* ``q8``     ``[rows, Kp]``    uint8 copy, K-major, ``Kp = ceil(D/128)*128``, zero padded   (2 KB / molecule;
  the TMA / tcgen05 ``kind::i8`` operand)
* ``rowsum`` / ``rowsq`` ``[rows]`` int32 exact row norms (popcount for bits)
* a classification (``binary`` / ``count`` / ``real``) that picks the kernel family.

Replaces the ``torch.sum(x**2)`` / ``torch.sum(x)`` reductions of the reference
(tanimoto_kernel.py:37-38, minmax_kernel.py:33-34).
"""
from __future__ import annotations

import weakref
from dataclasses import dataclass, field
from typing import Optional

import torch

from . import _lib

K_ALIGN = 128

_DTYPE_CODES = {
    torch.float32: _lib.GK_F32,
    torch.float64: _lib.GK_F64,
    torch.int64: _lib.GK_I64,
    torch.int32: _lib.GK_I32,
    torch.uint8: _lib.GK_U8,
    torch.float16: _lib.GK_F16,
    torch.bfloat16: _lib.GK_BF16,
    torch.int8: _lib.GK_I8,
    torch.int16: _lib.GK_I16,
    torch.bool: _lib.GK_BOOL,
}


def padded_k(dim: int) -> int:
    return max(K_ALIGN, (dim + K_ALIGN - 1) // K_ALIGN * K_ALIGN)


def stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


@dataclass
class PackedFingerprints:
    bits: torch.Tensor
    q8: torch.Tensor
    rowsum: torch.Tensor
    rowsq: torch.Tensor
    flags_dev: torch.Tensor
    rows: int
    dim: int
    kp: int
    _flags: Optional[int] = field(default=None, repr=False)
    _q4: Optional[torch.Tensor] = field(default=None, repr=False)
    _q4_owner: Optional["PackedFingerprints"] = field(default=None, repr=False)
    _row0: int = 0

    @property
    def device(self) -> torch.device:
        return self.q8.device

    @property
    def flags(self) -> int:
        """Classification bits; the first read synchronises on the pack kernel (4-byte D2H)."""
        if self._flags is None:
            self._flags = int(self.flags_dev.item())
        return self._flags

    @property
    def kind(self) -> str:
        f = self.flags
        if f == 0:
            return "binary"
        if not (f & _lib.GK_FLAG_NONU8):
            return "count"
        return "real"

    @property
    def k4_bytes(self) -> int:
        """Row length in bytes of the packed-e2m1 operand (padded to 256 fingerprint bits)."""
        return (self.kp + 255) // 256 * 128

    def q4(self) -> torch.Tensor:
        """``[rows, k4_bytes]`` packed e2m1 nibbles (1.0/0.0) for the kind::mxf4 Gram kernel, built
        lazily from ``bits`` by ``gk_expand_bits_e2m1`` and cached (row slices share the parent's)."""
        if self._q4_owner is not None:
            return self._q4_owner.q4()[self._row0:self._row0 + self.rows]
        if self._q4 is None:
            q4 = torch.empty((self.rows, self.k4_bytes), dtype=torch.uint8, device=self.device)
            if self.rows > 0:
                with torch.cuda.device(self.device):
                    rc = _lib.load().gk_expand_bits_e2m1(self.bits.data_ptr(), self.kp // 32, self.rows,
                                                         self.kp // 32, q4.data_ptr(), self.k4_bytes,
                                                         stream_ptr(self.device))
                _lib.check(rc, "gk_expand_bits_e2m1")
            self._q4 = q4
        return self._q4

    def rows_slice(self, start: int, stop: int) -> "PackedFingerprints":
        """Zero-copy view of a contiguous row range (row blocks / batch entries)."""
        owner = self._q4_owner if self._q4_owner is not None else self
        return PackedFingerprints(
            self.bits[start:stop], self.q8[start:stop], self.rowsum[start:stop], self.rowsq[start:stop],
            self.flags_dev, stop - start, self.dim, self.kp, self._flags,
            None, owner, self._row0 + start,
        )


def pack(x: torch.Tensor) -> PackedFingerprints:
    """Pack a 2-D CUDA tensor ``[rows, D]`` (any supported dtype) on its device's current stream."""
    if x.ndim != 2:
        raise ValueError(f"pack expects a 2-D tensor, got shape {tuple(x.shape)}")
    if not x.is_cuda:
        raise _lib.GaucheB200Error("pack expects a CUDA tensor (gauche_b200 has no CPU path)")
    if x.dtype not in _DTYPE_CODES:
        raise TypeError(f"unsupported fingerprint dtype {x.dtype}")
    if x.stride(-1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    rows, dim = x.shape
    kp = padded_k(dim)
    dev = x.device
    bits = torch.empty((rows, kp // 32), dtype=torch.int32, device=dev)
    q8 = torch.empty((rows, kp), dtype=torch.uint8, device=dev)
    rowsum = torch.empty((rows,), dtype=torch.int32, device=dev)
    rowsq = torch.empty((rows,), dtype=torch.int32, device=dev)
    flags = torch.zeros((1,), dtype=torch.int32, device=dev)
    if rows > 0 and dim == 0:
        for t in (bits, q8, rowsum, rowsq):
            t.zero_()
    elif rows > 0:
        lib = _lib.load()
        with torch.cuda.device(dev):
            rc = lib.gk_pack_classify(
                x.data_ptr(), _DTYPE_CODES[x.dtype], rows, dim, x.stride(0) if rows > 1 else max(dim, 1),
                bits.data_ptr(), kp // 32, q8.data_ptr(), kp, rowsum.data_ptr(), rowsq.data_ptr(),
                flags.data_ptr(), stream_ptr(dev),
            )
        _lib.check(rc, "gk_pack_classify")
    return PackedFingerprints(bits, q8, rowsum, rowsq, flags, rows, dim, kp)


# ------------------------------------------------------------------------------------------
# Operand cache.  TanimotoKernel/MinMaxKernel have no hyper-parameters, yet GPyTorch re-evaluates
# them on the same train tensor in every L-BFGS closure.  Entries are keyed on
# tensor identity (weak reference) + in-place version counter, so a recycled address can never hit.
# ------------------------------------------------------------------------------------------
class _PackCache:
    def __init__(self, max_entries: int = 8):
        self.max_entries = max_entries
        self._entries: dict[int, tuple] = {}
        self.hits = 0
        self.misses = 0
        self.enabled = True

    def get(self, x: torch.Tensor) -> PackedFingerprints:
        if not self.enabled:
            return pack(x)
        key = id(x)
        ent = self._entries.get(key)
        if ent is not None:
            ref, version, shape, stride, packed = ent
            if ref() is x and version == x._version and shape == tuple(x.shape) and stride == x.stride():
                self.hits += 1
                return packed
            del self._entries[key]
        self.misses += 1
        packed = pack(x)
        if len(self._entries) >= self.max_entries:
            self._entries.pop(next(iter(self._entries)))
        try:
            ref = weakref.ref(x, lambda _r, k=key, d=self._entries: d.pop(k, None))
        except TypeError:
            return packed
        self._entries[key] = (ref, x._version, tuple(x.shape), x.stride(), packed)
        return packed

    def clear(self) -> None:
        self._entries.clear()


pack_cache = _PackCache()
