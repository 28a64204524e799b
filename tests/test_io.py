"""Packed wire / on-disk format (gauche_b200/io.py): numpy-level round trips on the CPU, and — on the GPU — that a
memory-mapped library streamed through TanimotoScreener gives the scores of the dense tensors it stands for."""
import numpy as np
import pytest
import torch

from gauche_b200 import io


def _fragprints(rows, seed):
    rng = np.random.default_rng(seed)
    bits = (rng.random((rows, 2048)) < 0.03).astype(np.int64)
    counts = rng.poisson(0.5, size=(rows, 85)).astype(np.int64)
    return np.concatenate((bits, counts), axis=1)


def test_pack_roundtrip_bits_and_fragprints(tmp_path):
    x = _fragprints(53, 0)
    bits, counts = io.pack_fingerprints(x[:, :2048])
    assert counts is None and bits.shape == (53, 256) and bits.dtype == np.uint8
    assert np.array_equal(bits, np.packbits(x[:, :2048].astype(bool), axis=1, bitorder="little"))
    lib = io.PackedLibrary.from_dense(x, n_bits=2048)
    assert (lib.rows, lib.n_bits, lib.n_counts, lib.dim) == (53, 2048, 85, 2133)
    assert np.array_equal(lib.dense(np.int64), x)
    io.save_packed(str(tmp_path / "lib"), lib)
    assert io.exists(str(tmp_path / "lib"))
    again = io.load_packed(str(tmp_path / "lib"))
    assert isinstance(again.bits, np.memmap)
    assert np.array_equal(again.dense(np.float64), x.astype(np.float64))
    # shards partition the rows; blocks cover a shard
    spans = [again.shard(r, 4) for r in range(4)]
    assert sum(s.rows for s, _ in spans) == 53 and [st for _, st in spans] == [0, 14, 27, 40]
    assert np.array_equal(np.concatenate([b.dense() for _, b in spans[1][0].blocks(5)]), x[14:27].astype(np.float32))


def test_odd_bit_lengths_and_dtypes():
    rng = np.random.default_rng(1)
    for d in (1, 7, 37, 300, 2133):
        x = rng.random((9, d)) < 0.3
        for dt in (np.bool_, np.uint8, np.int64, np.float32, np.float64):
            lib = io.PackedLibrary.from_dense(x.astype(dt))
            assert lib.bits.shape == (9, (d + 7) // 8)
            assert np.array_equal(lib.dense(np.bool_), x)


def test_pack_rejects_what_the_format_cannot_hold():
    x = np.zeros((3, 16))
    x[1, 2] = 2
    with pytest.raises(ValueError, match="exactly 0 or 1"):
        io.pack_fingerprints(x)
    io.pack_fingerprints(x, n_bits=2)                        # the 2 sits in a count column now
    x[0, 5] = 256
    with pytest.raises(ValueError, match=r"\[0, 255\]"):
        io.pack_fingerprints(x, n_bits=2)
    x[0, 5] = 0.5
    with pytest.raises(ValueError, match="integers"):
        io.pack_fingerprints(x, n_bits=2)
    with pytest.raises(ValueError):
        io.pack_fingerprints(np.zeros(4))
    with pytest.raises(TypeError):
        io.PackedLibrary(np.zeros((2, 2), dtype=np.int32), 16)


def test_header_mismatch_is_detected(tmp_path):
    lib = io.PackedLibrary.from_dense(_fragprints(5, 2), 2048)
    stem = str(tmp_path / "l")
    io.save_packed(stem, lib)
    np.save(stem + ".bits.npy", lib.bits[:3])
    with pytest.raises((ValueError, TypeError)):
        io.load_packed(stem)


@pytest.mark.gpu
@pytest.mark.parametrize("fragprints", [False, True])
def test_screener_streams_a_memory_mapped_library(tmp_path, fragprints):
    from gauche_b200.screening import TanimotoScreener

    dev = torch.device("cuda:0")
    x_train, x_lib = _fragprints(700, 3), _fragprints(3001, 4)
    if not fragprints:
        x_train, x_lib = x_train[:, :2048], x_lib[:, :2048]
    io.save_packed(str(tmp_path / "lib"), io.PackedLibrary.from_dense(x_lib, 2048))
    lib = io.load_packed(str(tmp_path / "lib"))
    alpha = torch.randn(700, generator=torch.Generator().manual_seed(1))
    scr = TanimotoScreener(torch.from_numpy(x_train).float().to(dev), alpha.to(dev), bias=0.2, block_rows=1024)
    shard, start = lib.shard(1, 2)
    got = scr.scores(shard)
    ref = scr.scores(torch.from_numpy(x_lib[start:start + shard.rows]).float().to(dev))
    assert torch.equal(got, ref)
    assert scr.last_operand_kind == (0 if fragprints else 1)      # int8 kernel for fragprints, e2m1 bits otherwise
    if not fragprints:
        assert scr.host_bytes_copied == shard.rows * 256
