"""Tile-sharded data parallelism: one process per GPU, one gradient all-reduce per optimizer step.

Tiles are independent units: inference shards the tile list with no collective;
training replicates the model, shards the tiles and sums the gradients once per optimizer step
over a single flat fp32 buffer (SUM, not AVG: the reference accumulates un-normalised tile
gradients, trainer.py:70-79).
"""
import torch
import torch.distributed as dist


def shard_tiles(n_tiles: int, rank: int, world: int):
    """Contiguous block partition of the tile list (spatially adjacent tiles stay on one GPU)."""
    base, rem = divmod(n_tiles, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def shard_tiles_weighted(weights, rank: int, world: int):
    """Contiguous partition of the tile list with balanced total WEIGHT (e.g. candidate points per tile): rank r
    owns the tiles whose cumulative weight midpoint falls into the r-th of `world` equal slices.  Deterministic and
    identical on every rank (no communication); spatially adjacent tiles stay together."""
    total = float(sum(weights))
    if total <= 0:
        return shard_tiles(len(weights), rank, world)
    start, end, acc = None, None, 0.0
    for i, w in enumerate(weights):
        mid = acc + 0.5 * w
        owner = min(int(mid * world / total), world - 1)
        if owner == rank:
            start = i if start is None else start
            end = i + 1
        acc += w
    return range(start, end) if start is not None else range(0, 0)


class FlatGradients:
    """Re-homes every parameter's ``.grad`` into one contiguous buffer so that the whole model is
    reduced by ONE collective (NCCL picks NVLS / NVSwitch on a B200 box).

    The buffer only stays the gradients' home while nobody replaces ``p.grad``: ``optimizer.zero_grad()`` (the
    reference calls it, trainer.py:89) defaults to ``set_to_none=True`` and would silently detach every gradient
    from the buffer.  Use ``zero_()`` / ``zero_grad()`` of this class instead; ``all_reduce`` re-checks the
    homes (``verify``) and re-installs them, carrying over gradients that were accumulated elsewhere."""

    def __init__(self, module: torch.nn.Module):
        self.params = [p for p in module.parameters() if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        ref = self.params[0]
        self.flat = torch.zeros(total, dtype=ref.dtype, device=ref.device)
        self._views = []
        offset = 0
        for p in self.params:
            n = p.numel()
            # as_strided keeps the parameter's own (e.g. channels_last) strides
            view = self.flat[offset:offset + n].as_strided(p.shape, p.stride())
            self._views.append(view)
            p.grad = view
            offset += n

    def zero_(self):
        self.verify()
        self.flat.zero_()

    zero_grad = zero_

    def verify(self):
        """Every ``p.grad`` must still be its view of the flat buffer.  A gradient that was set to None is
        re-homed (its slice zeroed); one that was re-allocated elsewhere is copied in and re-homed.
        Returns the number of gradients that had to be repaired."""
        repaired = 0
        for p, view in zip(self.params, self._views):
            g = p.grad
            if g is not None and g.data_ptr() == view.data_ptr() and g.stride() == view.stride():
                continue
            with torch.no_grad():
                if g is None:
                    view.zero_()
                else:
                    view.copy_(g)
            p.grad = view
            repaired += 1
        return repaired

    def all_reduce(self, async_op=False):
        self.verify()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=async_op)
        return None

    @property
    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()
