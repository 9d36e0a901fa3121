"""world_size-2 CPU (gloo) tests of the tile-sharded data-parallel host logic (parallel.py)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tomosar2height_b200.parallel import FlatGradients, shard_tiles


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Conv2d(1, 2, 3))
        net[2].to(memory_format=torch.channels_last)
        flat = FlatGradients(net)
        # every rank owns a contiguous block of the 7 tiles; per-tile losses are summed un-normalised
        tiles = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
        flat.zero_()
        for t in shard_tiles(7, rank, world):
            net[0](tiles[t]).square().sum().backward()
            net[2](tiles[t].view(1, 1, 2, 3).repeat(1, 1, 2, 1)).sum().backward()
        flat.all_reduce()
        out[rank] = flat.flat.clone()
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_matches_single_process():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    # single-process reference over all tiles
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Conv2d(1, 2, 3))
    tiles = torch.arange(7 * 6, dtype=torch.float32).view(7, 6) / 10.0
    for t in range(7):
        net[0](tiles[t]).square().sum().backward()
        net[2](tiles[t].view(1, 1, 2, 3).repeat(1, 1, 2, 1)).sum().backward()
    want = torch.cat([p.grad.flatten() for p in net.parameters()])
    assert torch.allclose(out[0], out[1])           # replicas agree after the SUM all-reduce
    assert torch.allclose(out[0], want, rtol=1e-5, atol=1e-6)


def test_shard_tiles_partition():
    for n in (0, 1, 7, 32, 770):
        for world in (1, 2, 4, 8):
            parts = [list(shard_tiles(n, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))                       # contiguous, complete, disjoint
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_flat_gradients_are_views_of_one_buffer():
    net = torch.nn.Linear(4, 3)
    flat = FlatGradients(net)
    net(torch.ones(2, 4)).sum().backward()
    assert flat.flat.abs().sum() > 0
    assert net.weight.grad.data_ptr() == flat.flat.data_ptr()
    flat.zero_()
    assert float(net.weight.grad.abs().sum()) == 0.0 and flat.nbytes == 15 * 4
