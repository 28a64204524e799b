"""The path's only collective on real hardware: TanimotoScreener.screen() over NCCL.  Two ranks (one process per
GPU, spawned here) must return the same global top-k as a single-rank screen of the whole library.  Skipped with
fewer than two GPUs (the gloo twin, tests/test_screen_gloo.py, covers the host logic on CPU)."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_total: int, k: int, out_path: str) -> None:
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    from gauche_b200.screening import TanimotoScreener, shard_bounds
    from gauche_b200.synthetic import ecfp_bits

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    train = ecfp_bits(3000, 2048, seed=4, device=dev)
    alpha = torch.randn(3000, generator=torch.Generator().manual_seed(6)).to(dev)
    lib = ecfp_bits(n_total, 2048, seed=5, device=dev)          # same library on every rank; each screens its shard
    start, stop = shard_bounds(n_total, world, rank)
    scr = TanimotoScreener(train, alpha, bias=0.1, block_rows=4096)
    vals, idx = scr.screen(lib[start:stop], k, index_base=start)
    torch.cuda.synchronize()
    torch.save({"vals": vals.cpu(), "idx": idx.cpu()}, f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL)")
def test_two_rank_nccl_screen_equals_single_rank(tmp_path):
    import torch.multiprocessing as mp

    from gauche_b200.screening import TanimotoScreener
    from gauche_b200.synthetic import ecfp_bits

    n_total, k = 20_011, 257
    out = str(tmp_path / "topk")
    mp.spawn(_worker, args=(2, _free_port(), n_total, k, out), nprocs=2, join=True)
    r0, r1 = torch.load(out + ".0"), torch.load(out + ".1")
    assert torch.equal(r0["vals"], r1["vals"]) and torch.equal(r0["idx"], r1["idx"])      # identical on every rank
    dev = torch.device("cuda:0")
    train = ecfp_bits(3000, 2048, seed=4, device=dev)
    alpha = torch.randn(3000, generator=torch.Generator().manual_seed(6)).to(dev)
    lib = ecfp_bits(n_total, 2048, seed=5, device=dev)
    vals, idx = TanimotoScreener(train, alpha, bias=0.1, block_rows=4096).screen(lib, k)
    assert torch.equal(vals.cpu(), r0["vals"]) and torch.equal(idx.cpu(), r0["idx"])
