"""Packed wire / on-disk format of fingerprint libraries (numpy only — no torch, no CUDA, no RDKit).

The reference's featurisers hand dense arrays to the kernels: ``ecfp_fingerprints`` returns ``np.array(fps)``
(gauche/representations/fingerprints.py:72-93, one int64 per bit: 16 KB per 2048-bit molecule) and the
``ecfp_fragprints`` representation concatenates 2048 bits with 85 fragment counts
(gauche/dataloader/molprop_loader.py:122-144).  A million-compound library in that form is 16 GB; in the format
below it is 256 MB (+ 85 MB of counts) and is exactly what the tensor-core kernels read:

* ``bits``    ``uint8 [rows, ceil(n_bits/8)]``, little-endian bit order — bit ``i`` of byte ``b`` is fingerprint
  bit ``8b+i`` (``np.packbits(x, axis=1, bitorder="little")``); viewed as 32-bit words this is the operand
  ``gk_tanimoto_gram_bits`` / ``gk_tanimoto_rowdot(operand_kind=2)`` expand inside the SM
* ``counts``  ``uint8 [rows, n_counts]`` (fragprints only): the trailing count columns, values 0..255

On disk a library ``<stem>`` is ``<stem>.json`` (header) + ``<stem>.bits.npy`` (+ ``<stem>.counts.npy``), plain
``.npy`` files so that ``np.load(mmap_mode="r")`` gives zero-copy row shards to every rank.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass
from typing import Iterator, Optional, Tuple

import numpy as np

FORMAT = "gauche_b200.packed/1"


def pack_fingerprints(x: np.ndarray, n_bits: Optional[int] = None) -> Tuple[np.ndarray, Optional[np.ndarray]]:
    """Dense ``[rows, D]`` fingerprints (any integer / bool / float dtype holding integers) -> ``(bits, counts)``.

    ``n_bits`` = number of leading bit columns (default: all ``D``).  Those must be exactly 0/1 and are packed 8 per
    byte; the remaining ``D - n_bits`` columns (fragment counts of a fragprint) must be integers in ``[0, 255]`` and
    are kept as ``uint8``.  ``counts`` is ``None`` when there are no count columns.  Emitter-side twin of
    ``gk_host_pack_bits``: what an RDKit featuriser should write instead of ``np.array(fps)``."""
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError(f"pack_fingerprints expects [rows, D], got shape {x.shape}")
    d = x.shape[1]
    n_bits = d if n_bits is None else int(n_bits)
    if not 0 <= n_bits <= d:
        raise ValueError(f"n_bits = {n_bits} outside [0, {d}]")
    head, tail = x[:, :n_bits], x[:, n_bits:]
    if head.dtype != np.bool_ and not np.all((head == 0) | (head == 1)):
        raise ValueError("bit columns must be exactly 0 or 1")
    bits = np.packbits(head.astype(np.bool_), axis=1, bitorder="little")
    counts = None
    if tail.shape[1]:
        if tail.dtype.kind == "f" and not np.all(tail == np.floor(tail)):
            raise ValueError("count columns must hold integers")
        if tail.size and (tail.min() < 0 or tail.max() > 255):
            raise ValueError("count columns must lie in [0, 255]")
        counts = np.ascontiguousarray(tail.astype(np.uint8))
    return np.ascontiguousarray(bits), counts


def unpack_fingerprints(bits: np.ndarray, n_bits: int, counts: Optional[np.ndarray] = None,
                        dtype=np.float32) -> np.ndarray:
    """Inverse of ``pack_fingerprints``: the dense ``[rows, n_bits (+ n_counts)]`` matrix in ``dtype``."""
    dense = np.unpackbits(np.asarray(bits), axis=1, bitorder="little")[:, :n_bits].astype(dtype)
    if counts is not None:
        dense = np.concatenate((dense, np.asarray(counts).astype(dtype)), axis=1)
    return dense


@dataclass
class PackedLibrary:
    """A fingerprint library in the packed format (in memory or memory-mapped)."""

    bits: np.ndarray                      # uint8 [rows, ceil(n_bits / 8)]
    n_bits: int
    counts: Optional[np.ndarray] = None   # uint8 [rows, n_counts]

    def __post_init__(self):
        if self.bits.ndim != 2 or self.bits.dtype != np.uint8:
            raise TypeError("bits must be a 2-D uint8 array")
        if self.bits.shape[1] * 8 < self.n_bits:
            raise ValueError(f"{self.bits.shape[1]} bytes per row cannot hold {self.n_bits} bits")
        if self.counts is not None and (self.counts.ndim != 2 or self.counts.dtype != np.uint8
                                        or self.counts.shape[0] != self.bits.shape[0]):
            raise TypeError("counts must be a uint8 [rows, n_counts] array")

    @property
    def rows(self) -> int:
        return int(self.bits.shape[0])

    @property
    def n_counts(self) -> int:
        return 0 if self.counts is None else int(self.counts.shape[1])

    @property
    def dim(self) -> int:
        """Feature dimension of the dense matrix this library stands for."""
        return self.n_bits + self.n_counts

    def __len__(self) -> int:
        return self.rows

    def rows_slice(self, start: int, stop: int) -> "PackedLibrary":
        """Zero-copy view of a contiguous row range (a rank's shard, a streaming block)."""
        return PackedLibrary(self.bits[start:stop], self.n_bits, None if self.counts is None else self.counts[start:stop])

    def shard(self, rank: int, world_size: int) -> Tuple["PackedLibrary", int]:
        """This rank's contiguous row block (sizes differ by at most one) and its first global row index."""
        base, rem = divmod(self.rows, world_size)
        start = rank * base + min(rank, rem)
        stop = start + base + (1 if rank < rem else 0)
        return self.rows_slice(start, stop), start

    def blocks(self, block_rows: int) -> Iterator[Tuple[int, "PackedLibrary"]]:
        for r0 in range(0, self.rows, block_rows):
            yield r0, self.rows_slice(r0, min(self.rows, r0 + block_rows))

    def dense(self, dtype=np.float32) -> np.ndarray:
        return unpack_fingerprints(self.bits, self.n_bits, self.counts, dtype)

    @classmethod
    def from_dense(cls, x: np.ndarray, n_bits: Optional[int] = None) -> "PackedLibrary":
        bits, counts = pack_fingerprints(x, n_bits)
        return cls(bits, x.shape[1] if n_bits is None else int(n_bits), counts)


def save_packed(stem: str, lib: PackedLibrary) -> None:
    """Write ``<stem>.json`` + ``<stem>.bits.npy`` (+ ``<stem>.counts.npy``)."""
    np.save(stem + ".bits.npy", np.ascontiguousarray(lib.bits))
    if lib.counts is not None:
        np.save(stem + ".counts.npy", np.ascontiguousarray(lib.counts))
    header = {"format": FORMAT, "rows": lib.rows, "n_bits": lib.n_bits, "n_counts": lib.n_counts,
              "bitorder": "little", "bytes_per_row": int(lib.bits.shape[1])}
    with open(stem + ".json", "w") as f:
        json.dump(header, f)


def load_packed(stem: str, mmap: bool = True) -> PackedLibrary:
    """Open a library written by ``save_packed``; ``mmap=True`` maps the arrays read-only (rows are paged in as the
    screening loop touches them, so every rank can open the whole file and read only its shard)."""
    with open(stem + ".json") as f:
        header = json.load(f)
    if header.get("format") != FORMAT or header.get("bitorder") != "little":
        raise ValueError(f"{stem}.json is not a {FORMAT} header")
    mode = "r" if mmap else None
    bits = np.load(stem + ".bits.npy", mmap_mode=mode)
    counts = np.load(stem + ".counts.npy", mmap_mode=mode) if header["n_counts"] else None
    lib = PackedLibrary(bits, int(header["n_bits"]), counts)
    if lib.rows != header["rows"] or lib.n_counts != header["n_counts"] or bits.shape[1] != header["bytes_per_row"]:
        raise ValueError(f"{stem}: arrays do not match the header")
    return lib


def exists(stem: str) -> bool:
    return os.path.exists(stem + ".json") and os.path.exists(stem + ".bits.npy")
