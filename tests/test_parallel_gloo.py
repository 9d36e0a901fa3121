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


def test_flat_gradients_survive_zero_grad_set_to_none():
    """The reference trainer calls optimizer.zero_grad() (trainer.py:89), whose default sets every .grad to None:
    FlatGradients must notice and re-home the gradients instead of reducing a stale buffer."""
    net = torch.nn.Linear(4, 3)
    flat = FlatGradients(net)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    net(torch.ones(2, 4)).sum().backward()
    opt.zero_grad()                                   # set_to_none=True: grads leave the buffer
    assert net.weight.grad is None
    net(torch.ones(2, 4)).sum().backward()           # autograd allocates fresh gradients outside the buffer
    assert net.weight.grad.data_ptr() != flat.flat.data_ptr()
    want = torch.cat([p.grad.flatten().clone() for p in net.parameters()])
    assert flat.verify() == 2                         # both re-homed, values carried over
    assert net.weight.grad.data_ptr() == flat.flat.data_ptr()
    assert torch.equal(flat.flat, want)
    flat.all_reduce()                                 # single process: a no-op, but it verifies first
    assert flat.verify() == 0
    opt.zero_grad()
    flat.zero_()                                      # None grads are re-homed as zeros
    assert net.bias.grad is not None and float(flat.flat.abs().sum()) == 0.0


def _trainer_worker(rank, world, port, out):
    """Trainer.train_step accumulation (trainer.py:70-89) on CPU stand-ins: grads summed over ranks once per
    optimizer step, identical parameters afterwards."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tomosar2height_b200.trainer import Trainer

        class Toy(torch.nn.Module):   # same call signature / outputs as TomoSAR2Height.forward
            def __init__(self):
                super().__init__()
                torch.manual_seed(0)
                self.fc = torch.nn.Linear(3, 16)

            def forward(self, input_cloud=None, input_image=None):
                return self.fc(input_cloud.mean(1)).view(-1, 4, 4, 1), None

        model = Toy()
        opt = torch.optim.SGD(model.parameters(), lr=0.5)
        tr = Trainer(model, opt, micro_batch=2, use_cuda_graph=False, optimize_every=3)
        g = torch.Generator().manual_seed(rank)
        stepped_at = []
        for k in range(6):
            cloud = torch.rand(1, 10, 3, generator=g)
            loss, stepped = tr.train_step({"inputs": cloud, "dsm": torch.ones(1, 4, 4)})
            if stepped:
                stepped_at.append(k)
        out[rank] = (stepped_at, model.fc.weight.detach().clone())
    finally:
        dist.destroy_process_group()


def test_trainer_accumulates_and_reduces_once_per_step():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_trainer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0] == out[1][0] == [2, 5]              # an optimizer step every 3 tiles per rank
    assert torch.allclose(out[0][1], out[1][1])          # replicas stay identical: gradients were summed over ranks


def test_shard_tiles_weighted_balances_contiguous_blocks():
    from tomosar2height_b200.parallel import shard_tiles_weighted
    import random
    rnd = random.Random(0)
    for n in (1, 7, 770):
        weights = [rnd.choice([0, 1, 5, 40, 400]) for _ in range(n)]
        for world in (1, 2, 4, 8):
            parts = [list(shard_tiles_weighted(weights, r, world)) for r in range(world)]
            assert sum(parts, []) == list(range(n))                     # contiguous, complete, disjoint, in rank order
            if n == 770:
                loads = [sum(weights[i] for i in p) for p in parts]
                assert max(loads) <= sum(weights) / world + 400          # within one (heaviest) tile of the ideal share
    assert [list(shard_tiles_weighted([0, 0, 0], r, 2)) for r in range(2)] == [[0, 1], [2]]  # no weight: by count
